function [W, H, cost] = lnmf(V, num_basis_elems, config)
% LNMF  Drop-in for the toolbox's lnmf.m (same signature, lnmf.m:1) backed by libnmfb200.so
% (nmfb_lnmf: fused KL kernels, num_basis_elems <= 128).  As in the reference the cost vector is
% not trimmed when the loop stops early.  Not runnable in the build image (no MATLAB).
if nargin < 3, config = struct; end
[W, H, cost] = nmfb_mex('lnmf', single(V), num_basis_elems, config);
end
