function [S, F, T, P] = spectrogram(x, window, noverlap, nfft, fs, varargin)
% GNU Octave stand-in for the Signal Processing Toolbox call at radar_processing.m line 276 / 541:
%     [S, F, T, P] = spectrogram(x, window, noverlap, nfft, fs, 'yaxis')
% for a REAL vector x: one-sided, rows nfft/2+1 (nfft even) or (nfft+1)/2, F = (0:rows-1)*fs/nfft, columns
% k = fix((L - noverlap) / (length(window) - noverlap)), T = (length(window)/2 + (0:k-1)*hop) / fs, no padding, no detrend,
% P = |S|.^2 / (fs * sum(window.^2)) doubled except at DC and Nyquist.  ('yaxis' only affects plotting.)
% Only on the path for Octave (make_reference_golden.m adds octave_compat/ last); MATLAB uses its own.
    x = x(:);
    window = window(:);
    nwin = numel(window);
    hop = nwin - noverlap;
    L = numel(x);
    k = fix((L - noverlap) / hop);
    assert(k >= 1, 'spectrogram: signal shorter than the window');
    idx = bsxfun(@plus, (1:nwin)', (0:k-1) * hop);
    seg = bsxfun(@times, x(idx), window);
    X = fft(seg, nfft, 1);
    if mod(nfft, 2) == 0
        rows = nfft / 2 + 1;
    else
        rows = (nfft + 1) / 2;
    end
    S = X(1:rows, :);
    F = (0:rows-1)' * fs / nfft;
    T = (nwin / 2 + (0:k-1) * hop) / fs;
    P = abs(S) .^ 2 / (fs * sum(window .^ 2));
    if mod(nfft, 2) == 0
        P(2:end-1, :) = 2 * P(2:end-1, :);
    else
        P(2:end, :) = 2 * P(2:end, :);
    end
end
