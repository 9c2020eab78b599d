function [W, H, cost] = cnmf(V, num_basis_elems, context_len, config)
% CNMF  Drop-in for cnmf.m (cnmf.m:1) through libnmfb200.so: 'euclidean' / 'frobenius' on the Gram path,
% 'kl_divergence' | 'kl', 'is_divergence' | 'is' and 'ab_divergence' | 'ab' (config.alpha, config.beta;
% cnmf.m:137-147) on the two-weight path.  Multi-source cells with UNIFORM per-source settings are
% concatenated as in matlab/nmf.m (W is split along its second dimension); settings that differ between
% sources are not supported by the engine's cnmf and raise nmfb:unsupported.
if nargin < 4, config = struct; end
if ~iscell(num_basis_elems), num_basis_elems = {num_basis_elems}; end
sizes = cellfun(@(k) double(k), num_basis_elems(:)');
S = numel(sizes);
cfg = config;
is_H_cell = S > 1;
if isfield(cfg, 'H_init') && ~isempty(cfg.H_init)
    is_H_cell = iscell(cfg.H_init);
    if is_H_cell
        if numel(cfg.H_init) ~= S
            error(['Requested ', num2str(S), ' sources. Given ', num2str(numel(cfg.H_init)), ' initial encoding matrices.']);
        end
        cfg.H_init = cell2mat(cfg.H_init(:));
    end
end
is_W_cell = S > 1;
if isfield(cfg, 'W_init') && ~isempty(cfg.W_init)
    is_W_cell = iscell(cfg.W_init);
    if is_W_cell
        if numel(cfg.W_init) ~= S
            error(['Requested ', num2str(S), ' sources. Given ', num2str(numel(cfg.W_init)), ' initial basis tensors.']);
        end
        cfg.W_init = cat(2, cfg.W_init{:});                                    % m x sum(K_s) x T
    end
end
names = {'W_sparsity', 'H_sparsity', 'W_fixed', 'H_fixed'};
for f = 1 : numel(names)
    name = names{f};
    if ~isfield(cfg, name) || isempty(cfg.(name)) || ~iscell(cfg.(name)), continue; end
    v = cellfun(@(x) double(x), cfg.(name)(:)');
    if numel(v) > 1 && numel(v) ~= S
        error(['Requested ', num2str(S), ' sources. Given ', num2str(numel(v)), ' values for ', name, '.']);
    end
    if any(v ~= v(1))
        error('nmfb:unsupported', 'cnmf: per-source %s values are not supported by the accelerated path', name);
    end
    cfg.(name) = v(1);
end
[Wall, Hall, cost] = nmfb_mex('cnmf', single(V), sum(sizes), context_len, cfg);
if is_W_cell, W = mat2cell(Wall, size(Wall, 1), sizes, size(Wall, 3)); else, W = Wall; end
if is_H_cell, H = mat2cell(Hall, sizes, size(Hall, 2)); else, H = Hall; end
end
