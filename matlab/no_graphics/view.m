function view(varargin)
end
