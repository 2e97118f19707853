function m = jet(varargin)
    m = zeros(64, 3);
end
