function Upload(varargin)
% radar_processing_with_azure.m line 67 is a lost comment marker that MATLAB parses as the command
% Upload('JSON','to','Blob','Storage'); this no-op keeps the file byte-identical and runnable.
end
