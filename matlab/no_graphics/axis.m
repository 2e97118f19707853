function axis(varargin)
end
