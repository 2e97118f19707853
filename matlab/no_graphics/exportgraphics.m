function exportgraphics(fig, filename, varargin) %#ok<INUSL>
% no-graphics shim: an empty placeholder instead of spectrogram.png (line 344).
    fid = fopen(filename, 'w'); fclose(fid);
end
