function varargout = surf(varargin)
    if nargout, varargout{1} = []; end
end
