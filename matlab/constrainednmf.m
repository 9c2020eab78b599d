function [W, H, Z, A, cost] = constrainednmf(V, labels, num_basis_elems, config)
% CONSTRAINEDNMF  Drop-in for the toolbox's constrainednmf.m (constrainednmf.m:1) through libnmfb200.so.
% The label handling of the reference (lines 147-170: consecutive class numbers, unlabeled samples first,
% classes contiguous, the 0/1 matrix A) is host work and stays here, literally; the iteration loop
% (lines 183-258) runs on the GPU: W step as nmf.m, Z step on the class-summed gradients, H = Z*A.
% Extension: config.Z_init (num_basis_elems x (n_unlabeled + num_classes)), because the reference draws
% Z = rand(...) unconditionally (line 174).  Not runnable in the build image (no MATLAB) - see INTEGRATION.md.
if nargin < 4, config = struct; end
[~, n] = size(V);
assert(length(labels) == n, ['Length of the label vector not equal to number of samples. Length of label vector = ', ...
    num2str(length(labels)), '; number of samples = ', num2str(n)]);
labels = labels(:);
num_labeled_samps = length(find(labels > -1));                                   % line 149
[uniq_labels, ~, labels_processed] = unique(labels);                             % lines 151 / 156
if num_labeled_samps < n
    labels_processed = labels_processed - 1;
    labels_processed(labels_processed == 0) = -1;
    num_classes = length(uniq_labels) - 1;
else
    num_classes = length(uniq_labels);
end
[sorted_labels, sorted_idx] = sort(labels_processed, 'ascend');                   % line 163
n_unl = n - num_labeled_samps;
col2z = zeros(n, 1);                                                              % column of Z each ordered sample reads
col2z(1 : n_unl) = 0 : n_unl - 1;
col2z(n_unl + 1 : n) = n_unl + sorted_labels(n_unl + 1 : n) - 1;
nz = n_unl + num_classes;
cfg = config;
if isfield(cfg, 'Z_sparsity'), cfg.H_sparsity = cfg.Z_sparsity; cfg = rmfield(cfg, 'Z_sparsity'); end
if isfield(cfg, 'Z_fixed'), cfg.H_fixed = cfg.Z_fixed; cfg = rmfield(cfg, 'Z_fixed'); end
[W, H_sorted, Z, cost] = nmfb_mex('constrainednmf', single(V(:, sorted_idx)), num_basis_elems, int32(col2z), nz, cfg);
A_sorted = zeros(nz, n);                                                          % lines 166-170
A_sorted(sub2ind([nz, n], col2z' + 1, 1 : n)) = 1;
A = zeros(nz, n);                                                                 % lines 260-267: original sample order
A(:, sorted_idx) = A_sorted;
H = zeros(size(H_sorted), 'like', H_sorted);
H(:, sorted_idx) = H_sorted;
end
