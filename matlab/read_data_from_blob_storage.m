function info = read_data_from_blob_storage()
% Local-file stub (BASELINE north_star: blob I/O stubbed to local files): the two inputs must be in pwd.
    assert(exist('radar_data.xml', 'file') == 2 && exist('radar_data.raw.bin', 'file') == 2);
    info = struct('xml', 'radar_data.xml', 'raw', 'radar_data.raw.bin');
end
