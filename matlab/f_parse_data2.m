function [frame, frame_count, calib_data, sXML] = f_parse_data2(fdata)
% Shim of the vendor parser the reference calls at radar_processing.m line 86 but does not ship.
% Reads the FMCWRAW1 container defined in fmcw_radar_processing_b200/parse.py and returns what the UNTOUCHED
% reference loop expects: frame(k).Chirp = [NTS x PN x RX] complex double, ADC codes / 4095 (lines 199-202),
% calib_data = row vector [I_rx1 Q_rx1 I_rx2 Q_rx2 ...] / 4095 (lines 167-172), sXML with .Text leaves (line 94).
% f_parse_data2_raw.m is the variant the GPU path uses (frames stay int16).
    [raw, frame_count, calib_data, sXML] = f_parse_data2_raw(fdata);
    frame = frame_structs(raw);
end
