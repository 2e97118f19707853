function s = xml2struct(file)
% Minimal stand-in for File Exchange #28518 xml2struct (radar_processing.m lines 73-83 require it, the reference does
% not ship it): elements become nested struct fields, character data becomes the field 'Text'.  Written for the
% attribute-free XML that fmcw_radar_processing_b200/parse.py::write_recording emits (and the vendor files' subset the
% reference reads at lines 94-115); works in MATLAB and GNU Octave without Java.
    txt = fileread(file);
    txt = regexprep(txt, '<\?.*?\?>', '');
    txt = regexprep(txt, '<!--.*?-->', '');
    [s, rest] = parse_children(txt);
    assert(isempty(strtrim(rest)), 'xml2struct: trailing content');
end

function [node, txt] = parse_children(txt)
    node = struct();
    while true
        txt = strtrim_left(txt);
        if isempty(txt) || strncmp(txt, '</', 2)
            return;
        end
        if txt(1) ~= '<'
            k = find(txt == '<', 1);
            if isempty(k), k = numel(txt) + 1; end
            t = strtrim(txt(1:k-1));
            if ~isempty(t), node.Text = t; end
            txt = txt(k:end);
            continue;
        end
        tok = regexp(txt, '^<([A-Za-z_][\w\-\.:]*)([^>]*?)(/?)>', 'tokens', 'once');
        assert(~isempty(tok), 'xml2struct: malformed tag');
        name = tok{1};
        len = numel(regexp(txt, '^<[^>]*>', 'match', 'once'));
        txt = txt(len+1:end);
        if strcmp(tok{3}, '/')
            child = struct();
        else
            [child, txt] = parse_children(txt);
            close_tag = ['</' name '>'];
            txt = strtrim_left(txt);
            assert(strncmp(txt, close_tag, numel(close_tag)), ['xml2struct: missing ' close_tag]);
            txt = txt(numel(close_tag)+1:end);
        end
        fname = regexprep(name, '[^\w]', '_');
        if isfield(node, fname)
            if ~iscell(node.(fname)), node.(fname) = {node.(fname)}; end
            node.(fname){end+1} = child;
        else
            node.(fname) = child;
        end
    end
end

function t = strtrim_left(t)
    k = find(~isspace(t), 1);
    if isempty(k), t = ''; else, t = t(k:end); end
end
