function ylim(varargin)
end
