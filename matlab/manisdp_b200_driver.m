function [X, obj, data] = manisdp_b200_driver(kind, C, At, b, c, K, options)
% MANISDP_B200_DRIVER  Outer augmented-Lagrangian loop of the four primal ManiSDP drivers on top of the B200 engine.
%
% kind: 0 onlyunitdiag, 1 unitdiag, 2 unittrace, 3 general.  Everything n-sized (trust-region solve, KKT residues,
% eigen step, rank step, escape update) runs inside libmanisdp_b200.so through manisdp_mex; this file only sequences
% those calls and applies the scalar rules of the outer loop, which are the reference's:
%   stopping test / slow-progress abort  src/primal/ManiSDP_unitdiag.m:77-92 (and siblings)
%   sigma rule                            src/primal/ManiSDP_unitdiag.m:108-112
% Python twin of this file: manisdp_matlab_b200/solvers.py (that one is exercised by the test-suite).

names = {'onlyunitdiag', 'unitdiag', 'unittrace', 'general'};
dflt = defaults_of(kind);
f = fieldnames(dflt);
for i = 1:numel(f)
    if ~isfield(options, f{i}); options.(f{i}) = dflt.(f{i}); end
end
if ~isfield(options, 'seed'); options.seed = 0; end
if ~isfield(options, 'eig_tol'); options.eig_tol = 0; end
if ~isfield(options, 'use_graph'); options.use_graph = 1; end

fprintf('ManiSDP is starting...\n');
if kind == 0
    n = size(C, 1); m = n;
    if isfield(options, 'devices') && numel(options.devices) > 0
        % several GPUs from this one MATLAB session: a multi-GPU group (column-sharded trust-region solve)
        h = manisdp_mex('create', 0, n, sparse(C), double(options.devices(:)'));
    else
        h = manisdp_mex('create', 0, n, sparse(C));
    end
else
    n = K.s; m = length(b);
    h = manisdp_mex('create', kind, n, sparse(At), b, c);
    sigma = options.sigma0;
    manisdp_mex('set_dual', h, zeros(m, 1), sigma);
end
cleaner = onCleanup(@() manisdp_mex('destroy', h));
fprintf('SDP size: n = %i, m = %i\n', n, m);
layout = double(kind >= 2);          % unit-diag drivers hold Y as p x n, the other two as n x p
if isfield(options, 'Y0') && ~isempty(options.Y0)
    manisdp_mex('set_Y', h, options.Y0, layout);
else
    manisdp_mex('rand_Y', h, options.p0, options.seed);
end
tropts = struct('maxiter', options.TR_maxiter, 'maxinner', options.TR_maxinner, ...
                'tolgradnorm', options.tolgradnorm, 'use_graph', options.use_graph);
if kind == 1; every = 50; after = 100; else; every = 20; after = 50; end
data.status = 0; data.hv_count = 0; data.fac_size = [];
staged = false; gap0 = inf; pinf0 = inf; dinf0 = inf;
timespend = tic;
for iter = 1:options.AL_maxiter
    st = manisdp_mex('stats', h); p = st.p;
    data.fac_size(end+1) = p; %#ok<AGROW>
    if staged; manisdp_mex('line_search', h); end
    info = manisdp_mex('tr_solve', h, tropts);
    data.hv_count = data.hv_count + info.hv_count;
    gradnorm = info.gradnorm;
    eig_tol = options.eig_tol;
    if eig_tol <= 0; eig_tol = -options.tol; end   % adaptive eigen-step accuracy keyed to options.tol (manisdp_b200.h)
    k = manisdp_mex('kkt', h, options.delta, eig_tol, double(kind > 0));
    obj = k.obj; dinf = k.dinf; gap = k.gap; pinf = k.pinf;
    r = manisdp_mex('rank_cut', h, options.theta, 0);
    if kind == 0
        fprintf('Iter %d, obj:%0.8f, dinf:%0.1e, r:%d, p:%d, time:%0.2fs\n', iter, obj, dinf, r, p, toc(timespend));
        eta = dinf; worse = dinf > dinf0;
    else
        fprintf('Iter %d, obj:%0.8f, gap:%0.1e, pinf:%0.1e, dinf:%0.1e, gradnorm:%0.1e, r:%d, p:%d, sigma:%0.3f, time:%0.2fs\n', ...
                iter, obj, gap, pinf, dinf, gradnorm, r, p, sigma, toc(timespend));
        eta = max([gap, pinf, dinf]); worse = gap > gap0 && pinf > pinf0 && dinf > dinf0;
    end
    if eta < options.tol
        fprintf('Optimality is reached!\n');
        break;
    end
    if mod(iter, every) == 0
        if iter > after && worse
            data.status = 2;
            fprintf('Slow progress!\n');
            break;
        end
        gap0 = gap; pinf0 = pinf; dinf0 = dinf;
    end
    if iter == options.AL_maxiter; break; end
    if r <= p - 1; manisdp_mex('rank_cut', h, options.theta, 1); end
    nne = min(k.nneg, options.delta);
    if kind <= 1; nne = max(nne, 1); end
    staged = options.line_search == 1;
    manisdp_mex('escape', h, nne, options.alpha, options.line_search);
    if kind > 0
        if pinf < options.tau1*gradnorm
            sigma = max(sigma/options.gama, options.sigma_min);
        elseif pinf > options.tau2*gradnorm
            sigma = min(sigma*options.gama, options.sigma_max);
        end
        manisdp_mex('set_sigma', h, sigma);
    end
end
Y = manisdp_mex('get_Y', h, layout);
data.Y = Y;
if n <= 6000        % dense outputs as in the reference; larger problems return the factor only
    if layout == 0; X = Y'*Y; else; X = Y*Y'; end
    if kind == 0
        z = full(sum(C.*X)); S = C - diag(z); data.z = z;
    else
        y = manisdp_mex('get_dual', h); data.y = y;
        eS = reshape(c - At*y, n, n);
        if kind == 1; z = sum(X.*eS); S = eS - diag(z); data.z = z;
        elseif kind == 2; z = sum(eS.*X, 'all'); S = eS - z*speye(n); data.z = z;
        else; S = eS; end
    end
    data.X = X; data.S = S;
else
    X = [];
end
data.dinf = dinf; data.gradnorm = gradnorm; data.time = toc(timespend);
if kind > 0; data.gap = gap; data.pinf = pinf; end
if data.status == 0 && eta > options.tol
    data.status = 1;
    fprintf('Iteration maximum is reached!\n');
end
fprintf('ManiSDP: optimum = %0.8f, time = %0.2fs\n', obj, toc(timespend));
end

function d = defaults_of(kind)
% option defaults of the reference drivers (README L28-40 / 56-72 / 90-106 / 122-138)
switch kind
    case 0
        d = struct('p0', 2, 'AL_maxiter', 20, 'tol', 1e-8, 'theta', 1e-1, 'delta', 8, 'alpha', 0.5, ...
                   'tolgradnorm', 1e-8, 'TR_maxinner', 100, 'TR_maxiter', 40, 'line_search', 0);
    case 1
        d = struct('p0', 2, 'AL_maxiter', 300, 'gama', 2, 'sigma0', 1e-3, 'sigma_min', 1e-2, 'sigma_max', 1e7, ...
                   'tol', 1e-8, 'theta', 1e-3, 'delta', 8, 'alpha', 0.1, 'tolgradnorm', 1e-8, 'TR_maxinner', 20, ...
                   'TR_maxiter', 4, 'tau1', 1, 'tau2', 1, 'line_search', 0);
    case 2
        d = struct('p0', 1, 'AL_maxiter', 1000, 'gama', 2, 'sigma0', 1e1, 'sigma_min', 1e2, 'sigma_max', 1e7, ...
                   'tol', 1e-8, 'theta', 1e-2, 'delta', 8, 'alpha', 0.05, 'tolgradnorm', 1e-8, 'TR_maxinner', 40, ...
                   'TR_maxiter', 3, 'tau1', 1e-5, 'tau2', 1e-4, 'line_search', 1);
    otherwise
        d = struct('p0', 1, 'AL_maxiter', 1000, 'gama', 2, 'sigma0', 1e-2, 'sigma_min', 1e-1, 'sigma_max', 1e7, ...
                   'tol', 1e-8, 'theta', 1e-2, 'delta', 8, 'alpha', 0.1, 'tolgradnorm', 1e-8, 'TR_maxinner', 20, ...
                   'TR_maxiter', 4, 'tau1', 1e-2, 'tau2', 1e-1, 'line_search', 1);
end
end
