function [peak_idxs, peak_mags] = f_search_peak(sig, len, threshold, max_num, min_distance, max_distance, dist_per_bin)
% Shim of the range peak picker the reference calls at radar_processing.m lines 211/469 but does not ship.
% 5-point local maximum inside the distance gate, strongest first (ties: lowest index) -- the definition
% libfmcw_cuda and the Python oracle implement (peak_mode = 'strongest').
    cand = [];
    for n = 3:len-2
        fp = sig(n);
        rng = (n-1) * dist_per_bin;
        if rng >= min_distance && rng <= max_distance && fp >= threshold && ...
           fp >= sig(n-2) && fp >= sig(n-1) && fp > sig(n+1) && fp > sig(n+2)
            cand(end+1) = n; %#ok<AGROW>
        end
    end
    [~, order] = sortrows([-reshape(sig(cand), [], 1), cand(:)]);
    cand = cand(order');
    peak_idxs = cand(1:min(max_num, numel(cand)));
    peak_mags = reshape(sig(peak_idxs), 1, []);
end
