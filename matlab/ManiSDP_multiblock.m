function [X, obj, data] = ManiSDP_multiblock(At, b, c, K, options)
% Drop-in for src/primal/ManiSDP_multiblock.m:7 (SeDuMi data with several PSD blocks; the first K.nob blocks have a
% unit diagonal) on the B200 engine.  The blocks live on the device as one stacked factor (csrc/multiblock.cu): the
% trust-region solve, the per-block eig(S{i}), the per-block rank cut and the escape update run inside
% libmanisdp_b200.so through manisdp_mex; this file sequences the calls and applies the scalar rules of the outer loop
% (stopping test / slow-progress abort :98-113, sigma rule :154-158).  Python twin: solvers.py::ManiSDP_multiblock.
if nargin < 5; options = struct(); end
n = K.s(:)'; nb = length(n);
if ~isfield(K, 'nob'); K.nob = 0; end
d = struct('min_facsize', 2, 'AL_maxiter', 1000, 'gama', 2, 'sigma0', 1e-1, 'sigma_min', 1e-2, 'sigma_max', 1e7, ...
           'tol', 1e-8, 'theta', 1e-2, 'delta', 8, 'alpha', 0.1, 'tolgradnorm', 1e-8, 'TR_maxinner', 20, ...
           'TR_maxiter', 4, 'tau1', 1e1, 'tau2', 1e1, 'line_search', 0, 'seed', 0, 'use_graph', 1);
f = fieldnames(d);
for i = 1:numel(f)
    if ~isfield(options, f{i}); options.(f{i}) = d.(f{i}); end
end
if ~isfield(options, 'p0'); options.p0 = ones(nb, 1); end
fprintf('ManiSDP is starting...\n');
fprintf('SDP size: n = %i, m = %i\n', max(n), size(b, 1));
p = n;
for i = 1:nb
    if n(i) >= options.min_facsize; p(i) = options.p0(i); end
end
m = length(b);
h = manisdp_mex('create', 4, sum(n), sparse(At), b, c, double(n), double(K.nob));
cleaner = onCleanup(@() manisdp_mex('destroy', h));
sigma = options.sigma0;
manisdp_mex('set_dual', h, zeros(m, 1), sigma);
if isfield(options, 'Y0') && ~isempty(options.Y0)
    manisdp_mex('mb_set_Y', h, options.Y0);
else
    manisdp_mex('mb_rand_Y', h, double(p), options.seed);
end
tropts = struct('maxiter', options.TR_maxiter, 'maxinner', options.TR_maxinner, ...
                'tolgradnorm', options.tolgradnorm, 'use_graph', options.use_graph);
data.status = 0; data.hv_count = 0;
staged = false; gap0 = inf; pinf0 = inf; dinf0 = inf;
timespend = tic;
for iter = 1:options.AL_maxiter
    p = manisdp_mex('mb_widths', h, double(n));
    if staged; manisdp_mex('line_search', h); end
    info = manisdp_mex('tr_solve', h, tropts);
    data.hv_count = data.hv_count + info.hv_count;
    gradnorm = info.gradnorm;
    [k, dinfs] = manisdp_mex('mb_kkt', h, 1, double(n));
    obj = k.obj; gap = k.gap; pinf = k.pinf; dinf = k.dinf;
    fprintf('Iter %d, obj:%0.8f, gap:%0.1e, pinf:%0.1e, dinf:%0.1e, gradnorm:%0.1e, p_max:%d, sigma:%0.3f, time:%0.2fs\n', ...
            iter, obj, gap, pinf, dinf, gradnorm, max(p), sigma, toc(timespend));
    eta = max([gap, pinf, dinf]);
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
    if iter == options.AL_maxiter; break; end
    manisdp_mex('mb_update', h, options.theta, options.delta, options.alpha, options.line_search, ...
                options.min_facsize, double(n));
    staged = options.line_search == 1;
    if pinf < options.tau1*gradnorm
        sigma = max(sigma/options.gama, options.sigma_min);
    elseif pinf > options.tau2*gradnorm
        sigma = min(sigma*options.gama, options.sigma_max);
    end
    manisdp_mex('set_sigma', h, sigma);
end
Y = manisdp_mex('mb_get_Y', h, double(n));
y = manisdp_mex('get_dual', h);
cy = c - At*y;
X = cell(nb, 1); S = cell(nb, 1);
ind = 1;
for i = 1:nb
    X{i} = Y{i}'*Y{i};
    S{i} = reshape(cy(ind:ind+n(i)^2-1), n(i), n(i));
    if i <= K.nob; S{i} = S{i} - diag(sum(X{i}.*S{i})); end
    ind = ind + n(i)^2;
end
data.X = X; data.y = y; data.S = S; data.Y = Y; data.gap = gap; data.pinf = pinf; data.dinf = dinf;
data.dinfs = dinfs; data.gradnorm = gradnorm; data.time = toc(timespend);
if data.status == 0 && eta > options.tol
    data.status = 1;
    fprintf('Iteration maximum is reached!\n');
end
fprintf('ManiSDP: optimum = %0.8f, time = %0.2fs\n', obj, toc(timespend));
end
