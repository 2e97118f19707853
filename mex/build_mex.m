function build_mex()
% BUILD_MEX  Compiles the MEX gateway against libfmcw_cuda.so (MATLAB or GNU Octave).
%   Build the library first:  python -m fmcw_radar_processing_b200.build   (nvcc, sm_100a)
%   then, from MATLAB / Octave:  run('mex/build_mex.m')  ->  mex/fmcw_cuda_mex.<mexext>
    here = fileparts(mfilename('fullpath'));
    root = fileparts(here);
    lib = fullfile(root, 'fmcw_radar_processing_b200');
    inc = fullfile(root, 'include');
    src = fullfile(here, 'fmcw_cuda_mex.cpp');
    if exist('OCTAVE_VERSION', 'builtin')
        mkoctfile('--mex', ['-I' inc], ['-L' lib], '-lfmcw_cuda', ['-Wl,-rpath,' lib], '-o', fullfile(here, 'fmcw_cuda_mex'), src);
    else
        mex(['-I' inc], ['-L' lib], '-lfmcw_cuda', ['LDFLAGS=$LDFLAGS -Wl,-rpath,' lib], 'CXXFLAGS=$CXXFLAGS -std=c++17', ...
            '-outdir', here, src);
    end
    addpath(here); addpath(fullfile(root, 'matlab'));
    fprintf('built %s\n', fullfile(here, ['fmcw_cuda_mex.' mexext]));
end
