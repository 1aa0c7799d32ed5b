function [X, obj, data] = ManiSDP_unitdiag(At, b, c, K, options)
% Drop-in for src/primal/ManiSDP_unitdiag.m:7 (SeDuMi data, unit diagonal + affine constraints) on the B200 engine.
if nargin < 5; options = struct(); end
[X, obj, data] = manisdp_b200_driver(1, [], At, b, c, K, options);
end
