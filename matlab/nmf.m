function [W, H, cost] = nmf(V, num_basis_elems, config)
% NMF  Drop-in for the toolbox's nmf.m (same signature and cell-array conventions, nmf.m:1-65) backed by
% libnmfb200.so through the MEX gateway nmfb_mex.  All four divergences ('euclidean', 'kl_divergence' |
% 'kl', 'is_divergence' | 'is', 'ab_divergence' | 'ab' with config.alpha / config.beta) run on the GPU.
%
% Multi-source calls (num_basis_elems, W_init, H_init, W_sparsity, H_sparsity, W_fixed, H_fixed given as
% cell arrays, nmf.m:11-16) are handled here exactly as the reference's ValidateParameters does
% (nmf.m:270-401): the per-source loops of nmf.m:144-171 / 175-201 never refresh V_hat between sources,
% so S sources are ONE factorisation with the bases concatenated (W_all = cell2mat(W), H_all =
% cell2mat(H), nmf.m:136-137) and per-basis sparsity levels / fixed flags; the result is split back
% into cells (nmf.m:228-234).  Not runnable in the build image (no MATLAB) - see INTEGRATION.md.
if nargin < 3, config = struct; end
if ~iscell(num_basis_elems), num_basis_elems = {num_basis_elems}; end          % nmf.m:114-116
sizes = cellfun(@(k) double(k), num_basis_elems(:)');
S = numel(sizes);
cfg = config;

% initial factors (nmf.m:270-309): a cell must hold one matrix per source; a matrix means one source
is_H_cell = S > 1;
if isfield(cfg, 'H_init') && ~isempty(cfg.H_init)
    is_H_cell = iscell(cfg.H_init);
    if is_H_cell
        if numel(cfg.H_init) ~= S
            error(['Requested ', num2str(S), ' sources. Given ', num2str(numel(cfg.H_init)), ' initial encoding matrices.']);
        end
        cfg.H_init = cell2mat(cfg.H_init(:));                                  % {H_1; ...; H_S}
    end
end
is_W_cell = S > 1;
if isfield(cfg, 'W_init') && ~isempty(cfg.W_init)
    is_W_cell = iscell(cfg.W_init);
    if is_W_cell
        if numel(cfg.W_init) ~= S
            error(['Requested ', num2str(S), ' sources. Given ', num2str(numel(cfg.W_init)), ' initial basis matrices.']);
        end
        cfg.W_init = cell2mat(cfg.W_init(:)');                                 % {W_1 ... W_S}
    end
end

% per-source settings (nmf.m:311-401): one value is extended to all sources, S values become one value
% per basis column (the gateway's W_sparsity_k, H_sparsity_k, W_fixed_k, H_fixed_k)
what = struct('W_sparsity', 'sparsity levels', 'H_sparsity', 'sparsity levels', ...
              'W_fixed', 'update switches', 'H_fixed', 'update switches');
names = fieldnames(what);
for f = 1 : numel(names)
    name = names{f};
    if ~isfield(cfg, name) || isempty(cfg.(name)), continue; end
    v = cfg.(name);
    if iscell(v)
        if numel(v) > 1 && numel(v) ~= S
            error(['Requested ', num2str(S), ' sources. Given ', num2str(numel(v)), ' ', what.(name), '.']);
        end
        v = cellfun(@(x) double(x), v(:)');
    else
        v = double(v(1));
    end
    if numel(v) == 1
        cfg.(name) = v;
    else
        cfg.([name '_k']) = repelem(v, sizes);
        cfg = rmfield(cfg, name);
    end
end

[Wall, Hall, cost] = nmfb_mex('nmf', single(V), sum(sizes), cfg);

% nmf.m:228-234
if is_W_cell, W = mat2cell(Wall, size(Wall, 1), sizes); else, W = Wall; end
if is_H_cell, H = mat2cell(Hall, sizes, size(Hall, 2)); else, H = Hall; end
end
