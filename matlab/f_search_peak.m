function [peak_idxs, peak_mags] = f_search_peak(sig, len, threshold, max_num, min_distance, max_distance, dist_per_bin, peak_mode)
% Shim of the range peak picker the reference calls at radar_processing.m lines 211/469 but does not ship.
% 5-point local maximum inside the distance gate.  peak_mode (optional 8th argument, the reference passes seven):
%   'first' (default)  the first max_num peaks in increasing range -- the order the vendor SDK demo walks in;
%   'strongest'        the max_num largest peaks (ties: lowest index), what the comment at line 258 reads like.
% The same two definitions are what libfmcw_cuda (FMCW_PEAK_FIRST / FMCW_PEAK_STRONGEST) and the Python oracle implement.
    if nargin < 8, peak_mode = 'first'; end
    cand = [];
    for n = 3:len-2
        fp = sig(n);
        rng = (n-1) * dist_per_bin;
        if rng >= min_distance && rng <= max_distance && fp >= threshold && ...
           fp >= sig(n-2) && fp >= sig(n-1) && fp > sig(n+1) && fp > sig(n+2)
            cand(end+1) = n; %#ok<AGROW>
        end
    end
    if strcmp(peak_mode, 'strongest') && ~isempty(cand)
        [~, order] = sortrows([-reshape(sig(cand), [], 1), cand(:)]);
        cand = cand(order');
    end
    peak_idxs = cand(1:min(max_num, numel(cand)));
    peak_mags = reshape(sig(peak_idxs), 1, []);
end
