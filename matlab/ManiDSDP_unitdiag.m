function [X, obj, data] = ManiDSDP_unitdiag(A, b, c, K, options)
% Drop-in for src/dual/ManiDSDP_unitdiag.m:8 (the dual approach: Riemannian ADMM on the SOS form, unit-diagonal dual
% slack S = Y'*Y) on the B200 engine.  The closures of :171-191, the ADMM step of :71-88 and eig(X) run inside
% libmanisdp_b200.so (csrc/dual.cu) through manisdp_mex; this file sequences the calls and applies the scalar rules of
% the outer loop (stopping test / slow-progress abort :95-111, sigma rule :128-132).
% Python twin: manisdp_matlab_b200/solvers.py::ManiDSDP_unitdiag.
if nargin < 5; options = struct(); end
d = struct('ADMM_maxiter', 300, 'gama', 2, 'sigma0', 1e-3, 'sigma_min', 1e-3, 'sigma_max', 1e7, 'tol', 1e-8, ...
           'theta', 1e-3, 'delta', 8, 'alpha', 0.1, 'tolgradnorm', 1e-8, 'TR_maxinner', 20, 'TR_maxiter', 4, ...
           'tau1', 1e1, 'tau2', 1e2, 'line_search', 0, 'seed', 0, 'use_graph', 1, 'eig_tol', 0);
f = fieldnames(d);
for i = 1:numel(f)
    if ~isfield(options, f{i}); options.(f{i}) = d.(f{i}); end
end
if ~isfield(options, 'p0'); options.p0 = ceil(log(length(b))); end
if ~isfield(K, 'f'); K.f = 0; end
n = K.s; m = size(b, 1);
fprintf('ManiSDP is starting...\n');
fprintf('SDP size: n = %i, m = %i\n', n, m);
B = A(:, 1:K.f); Ap = A(:, K.f+1:end);
cf = c(1:K.f); cp = c(K.f+1:end);
if isfield(options, 'dAAt'); dAAt = full(options.dAAt(:)); else; dAAt = []; end
h = manisdp_mex('create', 5, n, sparse(Ap'), full(b), full(cp), dAAt, sparse(B), full(cf));
cleaner = onCleanup(@() manisdp_mex('destroy', h));
sigma = options.sigma0;
manisdp_mex('set_sigma', h, sigma);
if isfield(options, 'Y0') && ~isempty(options.Y0)
    manisdp_mex('set_Y', h, options.Y0, 0);
else
    manisdp_mex('rand_Y', h, options.p0, options.seed);
end
tropts = struct('maxiter', options.TR_maxiter, 'maxinner', options.TR_maxinner, ...
                'tolgradnorm', options.tolgradnorm, 'use_graph', options.use_graph);
data.status = 0; data.hv_count = 0; data.fac_size = []; data.seta = [];
staged = false; gap0 = inf; pinf0 = inf; dinf0 = inf;
timespend = tic;
for iter = 1:options.ADMM_maxiter
    st = manisdp_mex('stats', h); p = st.p;
    data.fac_size(end+1) = p; %#ok<AGROW>
    if staged; manisdp_mex('line_search', h); end
    info = manisdp_mex('tr_solve', h, tropts);
    data.hv_count = data.hv_count + info.hv_count;
    gradnorm = info.gradnorm;
    eig_tol = options.eig_tol;
    if eig_tol <= 0; eig_tol = -options.tol; end
    k = manisdp_mex('kkt', h, options.delta, eig_tol, 1);
    obj = k.obj; gap = k.gap; pinf = k.pinf; dinf = k.dinf;
    r = manisdp_mex('rank_cut', h, options.theta, 0);
    fprintf('Iter %d, obj:%0.8f, gap:%0.1e, pinf:%0.1e, dinf:%0.1e, gradnorm:%0.1e, r:%d, p:%d, sigma:%0.3f, time:%0.2fs\n', ...
            iter, obj, gap, pinf, dinf, gradnorm, r, p, sigma, toc(timespend));
    eta = max([gap, pinf, dinf]);
    data.seta(end+1) = eta; %#ok<AGROW>
    if eta < options.tol
        fprintf('Optimality is reached!\n');
        break;
    end
    if mod(iter, 50) == 0
        if iter > 100 && gap > gap0 && pinf > pinf0 && dinf > dinf0
            data.status = 2;
            fprintf('Slow progress!\n');
            break;
        end
        gap0 = gap; pinf0 = pinf; dinf0 = dinf;
    end
    if iter == options.ADMM_maxiter; break; end
    if r <= p - 1; manisdp_mex('rank_cut', h, options.theta, 1); end
    nne = max(min(k.nneg, options.delta), 1);
    staged = options.line_search == 1;
    manisdp_mex('escape', h, nne, options.alpha, options.line_search);
    if pinf < options.tau1*gradnorm
        sigma = max(sigma/options.gama, options.sigma_min);
    elseif pinf > options.tau2*gradnorm
        sigma = min(sigma*options.gama, options.sigma_max);
    end
    manisdp_mex('set_sigma', h, sigma);
end
Y = manisdp_mex('get_Y', h, 0);
y = manisdp_mex('get_dual', h);
[x, w] = manisdp_mex('dual_state', h, K.f);
S = Y'*Y;
if isempty(dAAt); dAAt = full(diag(Ap*Ap')); end
eX = reshape(x + Ap'*(b./dAAt), n, n);
X = eX - diag(sum(S.*eX));
data.X = X; data.y = y; data.S = S; data.w = w; data.Y = Y;
data.gap = gap; data.pinf = pinf; data.dinf = dinf; data.gradnorm = gradnorm; data.time = toc(timespend);
if data.status == 0 && eta > options.tol
    data.status = 1;
    fprintf('Iteration maximum is reached!\n');
end
fprintf('ManiDSDP: optimum = %0.8f, time = %0.2fs\n', obj, toc(timespend));
end
