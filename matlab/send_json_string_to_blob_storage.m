function send_json_string_to_blob_storage(filename) %#ok<INUSD>
% Local-file stub: the JSON stays in pwd.
end
