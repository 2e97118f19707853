function colormap(varargin)
end
