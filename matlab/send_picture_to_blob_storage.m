function send_picture_to_blob_storage(filename) %#ok<INUSD>
% Local-file stub: the PNG stays in pwd.
end
