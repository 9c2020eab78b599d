function [W, H, cost] = cnmf(V, num_basis_elems, context_len, config)
% CNMF  Drop-in for cnmf.m (cnmf.m:1): 'euclidean' / 'frobenius' on the GPU through libnmfb200.so.
if nargin < 4, config = struct; end
if iscell(num_basis_elems) && numel(num_basis_elems) == 1, num_basis_elems = num_basis_elems{1}; end
[W, H, cost] = nmfb_mex('cnmf', single(V), num_basis_elems, context_len, config);
end
