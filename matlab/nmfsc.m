function [W, H, cost] = nmfsc(V, num_basis_elems, config)
% NMFSC  Drop-in for nmfsc.m (nmfsc.m:1) through libnmfb200.so.
if nargin < 3, config = struct; end
[W, H, cost] = nmfb_mex('nmfsc', single(V), num_basis_elems, config);
end
