function V_hat = ReconstructFromDecomposition(W, H)
% Drop-in for ReconstructFromDecomposition.m:1 through libnmfb200.so.
if iscell(W), W = cell2mat(W); end
if iscell(H), H = cell2mat(H); end
V_hat = nmfb_mex('reconstruct', single(W), single(H));
end
