function regen_in_matlab(toolbox_dir)
% REGEN_IN_MATLAB  Run the REAL colinvaz/nmf-toolbox functions (nmf.m, cnmf.m, nmfsc.m, cnmfsc.m, lnmf.m)
% on the seeded inputs of every golden case and store their outputs, so that the NumPy oracle
% (oracle/nmf_oracle.py) - and through it every GPU parity test of this repository - can be pinned
% against the reference itself.  The build image has no MATLAB / Octave, which is why this step is left
% to whoever has one (see export_inputs_for_matlab.py for the three-step recipe).
%
%   regen_in_matlab('/path/to/nmf-toolbox')     % run from tests/golden
if nargin < 1, toolbox_dir = '.'; end
addpath(toolbox_dir);
io = fullfile(fileparts(mfilename('fullpath')), 'matlab_io');
files = dir(fullfile(io, '*_in.mat'));
for f = 1 : numel(files)
    in = load(fullfile(io, files(f).name));
    name = files(f).name(1 : end - 7);
    config = struct('W_init', in.W_init, 'H_init', in.H_init, 'maxiter', in.maxiter, 'tolerance', in.tolerance);
    opt = {'divergence', 'alpha', 'beta', 'W_sparsity', 'H_sparsity'};
    for k = 1 : numel(opt)
        if isfield(in, opt{k}), config.(opt{k}) = in.(opt{k}); end
    end
    tic;
    switch in.alg
        case 'nmf',    [W, H, cost] = nmf(in.V, in.K, config);
        case 'lnmf',   [W, H, cost] = lnmf(in.V, in.K, config);
        case 'cnmf',   [W, H, cost] = cnmf(in.V, in.K, in.T, config);
        case 'cnmfsc', [W, H, cost] = cnmfsc(in.V, in.K, in.T, config);
        case 'nmfsc',  [W, H, cost] = nmfsc(in.V, in.K, config);
        otherwise, error('unknown algorithm %s', in.alg);
    end
    seconds = toc;
    V_hat = ReconstructFromDecomposition(W, H);
    vhat_norm = norm(V_hat, 'fro');
    save(fullfile(io, [name '_out.mat']), 'W', 'H', 'cost', 'vhat_norm', 'seconds', '-v7');
    fprintf('%s: %d cost entries, cost(end) = %.10g, %.1f s\n', name, numel(cost), cost(end), seconds);
end
end
