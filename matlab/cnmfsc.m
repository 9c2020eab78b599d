function [W, H, cost] = cnmfsc(V, num_basis_elems, context_len, config)
% CNMFSC  Drop-in for the toolbox's cnmfsc.m (same signature, cnmfsc.m:1) backed by libnmfb200.so
% (nmfb_cnmfsc).  config.W_sparsity > 0 is not accelerated: call the original for that case (its W
% line search ends by step-size underflow, see include/nmfb200.h).  Not runnable in the build image.
if nargin < 4, config = struct; end
[W, H, cost] = nmfb_mex('cnmfsc', single(V), num_basis_elems, context_len, config);
end
