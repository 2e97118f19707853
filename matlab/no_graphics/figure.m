function h = figure(varargin)
% no-graphics shim (headless golden runs): the plotting block at radar_processing.m lines 332-348 becomes a no-op.
    h = 1;
end
