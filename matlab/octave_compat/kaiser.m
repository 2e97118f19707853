function w = kaiser(N, beta)
% Fallback for Octave without the signal package: symmetric Kaiser window, w(n) = I0(beta*sqrt(1-((n-a)/a)^2))/I0(beta).
    if nargin < 2, beta = 0.5; end
    if N == 1, w = 1; return; end
    a = (N - 1) / 2;
    n = (0:N-1)';
    w = besseli(0, beta * sqrt(max(0, 1 - ((n - a) / a) .^ 2))) / besseli(0, beta);
end
