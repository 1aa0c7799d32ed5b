% RUN_REFERENCE  Times the UNMODIFIED reference (wangjie212/ManiSDP-matlab + vendored Manopt 7.0) on the GPU box's host
% cores for the same instances bench.py / tools/run_configs.py use, so a MATLAB-equipped user can put the true CPU
% baseline next to the engine's numbers (the build container has no MATLAB; bench.py times the oracle port instead).
%   refroot = '/path/to/ManiSDP-matlab'; run_reference
addpath(genpath(refroot));
fprintf('maxNumCompThreads = %d\n', maxNumCompThreads);
% config 1: G-set G1 through ManiSDP_onlyunitdiag (example/example_maxcut.m:10-34)
fid = fopen(fullfile(refroot, 'data', 'Gset', 'G1.txt'), 'r'); hdr = fscanf(fid, '%d', 2);
E = fscanf(fid, '%f', [3, hdr(2)])'; fclose(fid);
n = hdr(1); A = sparse(E(:,1), E(:,2), E(:,3), n, n); A = A + A';
C = -(spdiags(sum(A, 2), 0, n, n) - A)/4;
rng(0); opts = struct('p0', 40);
tic; [~, fval, data] = ManiSDP_onlyunitdiag(C, opts); t = toc;
fprintf('G1: optimum %.8f, dinf %.1e, time %.2f s\n', fval, data.dinf, t);
% config 2: BQP q=60 through ManiSDP_unitdiag (example/example_bqp.m:31-41)
Q = load(fullfile(refroot, 'data', 'bqp_Q_60_1.txt')); e = load(fullfile(refroot, 'data', 'bqp_e_60_1.txt'));
[At, b, c, K] = bqpmom(60, Q, e); c = c/max(abs(c));
rng(0); tic; [~, fval, data] = ManiSDP_unitdiag(At, b, c, K, struct('tol', 1e-8)); t = toc;
fprintf('BQP-60: optimum %.8f, eta %.1e, time %.2f s\n', fval, max([data.gap, data.pinf, data.dinf]), t);
