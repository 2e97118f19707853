function clim(lims)
% clim (radar_processing.m line 337) was introduced in R2022a; older MATLAB and Octave call it caxis.
    caxis(lims);
end
