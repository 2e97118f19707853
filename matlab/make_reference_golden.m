function make_reference_golden(reference_dir, cases_dir, out_dir, no_graphics)
% MAKE_REFERENCE_GOLDEN  Runs the UNTOUCHED reference on the committed synthetic recordings and keeps what it writes.
%
%   make_reference_golden('/path/to/fmcw_radar_processing/radar-etl-pipeline')      % MATLAB or GNU Octave (>= 7: jsonencode)
%   octave --eval "addpath('matlab'); make_reference_golden('/path/to/radar-etl-pipeline')"
%
% reference_dir  folder holding the reference's radar_processing.m and radar_processing_with_azure.m (never copied here)
% cases_dir      default tests/golden/ref_cases : one folder per case with radar_data.xml + radar_data.raw.bin
%                (written by tests/golden/make_ref_cases.py from the seeded synthetic scenes)
% out_dir        default tests/golden/ref_out   : per case and branch, the JSON files the reference wrote
% no_graphics    default true: the surf/exportgraphics block (lines 332-348) is shimmed to no-ops (headless hosts)
%
% For every case the server entry point (main(input), radar_processing_with_azure.m line 9) is called with
% processAnimalActivity = 'no' and 'yes' in a scratch folder that holds the two input files.  The functions the
% reference calls but does not ship come from this folder (f_parse_data2, f_search_peak, xml2struct, Upload and the
% local-file stubs of the three blob functions); spectrogram / exportgraphics / clim / kaiser / chebwin stand-ins are
% added at the END of the path, so a MATLAB installation or Octave's signal package takes precedence.
% tests/test_reference_golden.py then pins the Python oracle (and, on a GPU box, libfmcw_cuda) to these files.
    here = fileparts(mfilename('fullpath'));
    root = fileparts(here);
    if nargin < 2 || isempty(cases_dir), cases_dir = fullfile(root, 'tests', 'golden', 'ref_cases'); end
    if nargin < 3 || isempty(out_dir), out_dir = fullfile(root, 'tests', 'golden', 'ref_out'); end
    if nargin < 4, no_graphics = true; end
    is_octave = exist('OCTAVE_VERSION', 'builtin') ~= 0;
    if is_octave
        try, pkg load signal; catch, disp('signal package not found: using matlab/octave_compat fallbacks'); end
    end
    addpath(here);                                          % shims of the unshipped functions + blob stubs
    addpath(reference_dir);                                 % the untouched reference
    if no_graphics, addpath(fullfile(here, 'no_graphics')); end
    if is_octave || ~exist('spectrogram', 'file')
        addpath(fullfile(here, 'octave_compat'), '-end');
    end
    d = dir(cases_dir);
    for i = 1:numel(d)
        if ~d(i).isdir || d(i).name(1) == '.', continue; end
        for flag = {'no', 'yes'}
            work = tempname(); mkdir(work); mkdir(fullfile(work, 'lib'));
            copyfile(fullfile(cases_dir, d(i).name, 'radar_data.xml'), work);
            copyfile(fullfile(cases_dir, d(i).name, 'radar_data.raw.bin'), work);
            old = cd(work);
            input = struct('processAnimalActivity', flag{1});
            try
                % the server entry point: `function result = main(input)` at line 9 of radar_processing_with_azure.m; MATLAB and
                % Octave dispatch on the FILE name
                result = radar_processing_with_azure(input);
                status = result.status;
            catch err
                status = ['exception: ' err.message];
            end
            cd(old);
            dst = fullfile(out_dir, d(i).name, flag{1});
            if ~exist(dst, 'dir'), mkdir(dst); end
            js = dir(fullfile(work, '*.json'));
            for k = 1:numel(js), copyfile(fullfile(work, js(k).name), dst); end
            fid = fopen(fullfile(dst, 'status.txt'), 'w'); fprintf(fid, '%s\n%s\n', status, version()); fclose(fid);
            fprintf('%s / %s: %s, %d JSON files\n', d(i).name, flag{1}, status, numel(js));
        end
    end
end
