function w = chebwin(M, at)
% Fallback for Octave without the signal package: Dolph-Chebyshev window, sidelobes at dB down (default 100, the value
% MATLAB uses when radar_processing.m line 139 calls chebwin(PN)); the same construction as scipy.signal.windows.chebwin.
    if nargin < 2, at = 100; end
    if M == 1, w = 1; return; end
    order = M - 1;
    beta = cosh(acosh(10 ^ (abs(at) / 20)) / order);
    k = (0:M-1)';
    x = beta * cos(pi * k / M);
    p = zeros(M, 1);
    p(x > 1) = cosh(order * acosh(x(x > 1)));
    p(x < -1) = (2 * mod(M, 2) - 1) * cosh(order * acosh(-x(x < -1)));
    in = abs(x) <= 1;
    p(in) = cos(order * acos(x(in)));
    if mod(M, 2)
        W = real(fft(p));
        n = (M + 1) / 2;
        w = [W(n:-1:2); W(1:n)];
    else
        p = p .* exp(1i * pi / M * k);
        W = real(fft(p));
        n = M / 2 + 1;
        w = [W(n:-1:2); W(2:n)];
    end
    w = w / max(w);
end
