function colorbar(varargin)
end
