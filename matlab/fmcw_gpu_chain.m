function r = fmcw_gpu_chain(cmd, frame_i16, calib_data, c)
% FMCW_GPU_CHAIN  Calls libfmcw_cuda through the MEX gateway in place of the per-frame loop
% (radar_processing.m lines 197-261) and the STFT block (lines 265-299 / 538-566).
%
%   r = fmcw_gpu_chain('run', frame_i16, calib_data, c)
%
% frame_i16  int16 [2 x NTS x PN x RX x N] ADC codes as returned by f_parse_data2_raw.m in this folder
% c          struct with the fmcw_configurations field names (radar_processing.m lines 645-672)
%
% r has the variables the rest of radar_processing.m uses, shaped exactly like the reference's:
%   range_tx1rx1_max_abs [256 x N], target_measurements (.strength/.range/.speed with the growing-matrix
%   shape of lines 245-250), T, log_freq_bins, interp_intensity [1024 x ncol], detected, tgt_range_idx,
%   tgt_doppler_idx.
% UNTESTED HERE: no MATLAB/Octave exists in the build environment (see INTEGRATION.md).
    o = fmcw_cuda_mex(cmd, frame_i16, calib_data, c);
    N = numel(o.detected);
    r.detected = logical(o.detected);
    r.tgt_range_idx = double(o.range_idx);            % 1-based
    r.tgt_doppler_idx = double(o.doppler_idx);        % 1-based, 9 = no Doppler peak (line 237)
    r.range_tx1rx1_max_abs = double(o.range_max_abs); % line 265
    % lines 157-159 allocate max_num_targets x frame_count, lines 245-250 write (fr_idx, j)
    s = zeros(c.max_num_targets, N); rg = s; sp = s;
    for fr = find(r.detected)
        s(fr, 1)  = o.range_mag(fr);
        rg(fr, 1) = (r.tgt_range_idx(fr) - 1) * c.dist_per_bin;
        sp(fr, 1) = (r.tgt_doppler_idx(fr) - c.Doppler_fft_size/2 - 1) * -c.fD_per_bin * c.Hz_to_mps_constant;
    end
    r.target_measurements = struct('strength', s, 'range', rg, 'speed', sp);
    if isfield(o, 'intensity') && ~isempty(o.intensity)
        r.T = o.T; r.log_freq_bins = o.frequency; r.interp_intensity = o.intensity;   % lines 276, 296, 299
    end
    r.slow_time_mag = o.slow_time_mag;
end
