function frame = frame_structs(frame_i16)
% frame(k).Chirp = [NTS x PN x RX] complex double, normalised by 4095 (input of the untouched reference loop).
    N = size(frame_i16, 5);
    frame = struct('Chirp', cell(1, N));
    for k = 1:N
        z = double(frame_i16(:, :, :, :, k)) / 4095;
        frame(k).Chirp = squeeze(complex(z(1, :, :, :), z(2, :, :, :)));
    end
end
