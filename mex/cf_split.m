function [indC,indF,SubG] = cf_split(S)
% Drop-in for the reference's AMG/cf_split.m: the C/F split on the GPU, the graph object in MATLAB.
[indC,indF] = cf_split_mex(S);
SubG = graph(S);            % AMG/cf_split.m:6
end
