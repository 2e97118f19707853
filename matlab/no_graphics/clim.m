function clim(varargin)
end
