function [W, H, cost] = cnmfsc(V, num_basis_elems, context_len, config)
% CNMFSC  Drop-in for the toolbox's cnmfsc.m (same signature, cnmfsc.m:1) backed by libnmfb200.so
% (nmfb_cnmfsc).  config.W_sparsity > 0 behaves as in the original (including its early "Algorithm
% converged" return, see include/nmfb200.h).  Not runnable in the build image.
if nargin < 4, config = struct; end
[W, H, cost] = nmfb_mex('cnmfsc', single(V), num_basis_elems, context_len, config);
end
