function [frame, frame_count, calib_data, sXML] = f_parse_data2_raw(fdata)
% GPU-path parser: like f_parse_data2.m (the shim of the vendor parser called at radar_processing.m line 86), but
% the frames stay raw.
% Reads the FMCWRAW1 container defined in fmcw_radar_processing_b200/parse.py.  For the GPU path `frame`
% is the raw int16 block [2 x NTS x PN x RX x N]; frame_structs() below converts it to the
% frame(k).Chirp [NTS x PN x RX] complex doubles the untouched reference loop expects.
    sXML = xml2struct([fdata '.xml']);
    fid = fopen([fdata '.raw.bin'], 'r', 'ieee-le');
    magic = fread(fid, 8, 'uint8=>char')';
    assert(strcmp(magic, 'FMCWRAW1'));
    hdr = fread(fid, 5, 'uint32');            % n_frames, n_rx, PN, NTS, n_cal
    frame_count = hdr(1);
    calib_data = fread(fid, 2*hdr(2)*hdr(5), 'double')' / 4095;
    frame = reshape(fread(fid, inf, 'int16=>int16'), [2, hdr(4), hdr(3), hdr(2), hdr(1)]);
    fclose(fid);
end
