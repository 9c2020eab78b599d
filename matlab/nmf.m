function [W, H, cost] = nmf(V, num_basis_elems, config)
% NMF  Drop-in for the toolbox's nmf.m (same signature, nmf.m:1) backed by libnmfb200.so.
% Single-source inputs with any of the four divergences ('euclidean', 'kl', 'is', 'ab' with
% config.alpha / config.beta) run on the GPU; cell-array sources with different per-source
% settings should call the original implementation.  Not runnable in the build image (no MATLAB) - see INTEGRATION.md.
if nargin < 3, config = struct; end
if iscell(num_basis_elems) && numel(num_basis_elems) == 1, num_basis_elems = num_basis_elems{1}; end
[W, H, cost] = nmfb_mex('nmf', single(V), num_basis_elems, config);
end
