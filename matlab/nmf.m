function [W, H, cost] = nmf(V, num_basis_elems, config)
% NMF  Drop-in for the toolbox's nmf.m (same signature, nmf.m:1) backed by libnmfb200.so.
% Single-source inputs with any of the four divergences ('euclidean', 'kl', 'is', 'ab' with
% config.alpha / config.beta) run on the GPU.  Cell-array (multi-source) calls: concatenate the
% sources' W_init / H_init and pass per-source sparsity levels / fixed flags as the per-basis
% vectors config.W_sparsity_k, H_sparsity_k, W_fixed_k, H_fixed_k (one entry per basis column,
% repelem(setting, num_basis_elems)); the result is split back with mat2cell.  Not runnable in the build image (no MATLAB) - see INTEGRATION.md.
if nargin < 3, config = struct; end
if iscell(num_basis_elems) && numel(num_basis_elems) == 1, num_basis_elems = num_basis_elems{1}; end
[W, H, cost] = nmfb_mex('nmf', single(V), num_basis_elems, config);
end
