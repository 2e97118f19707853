function close(varargin)
end
