function [X, obj, data] = ManiSDP(At, b, c, K, options)
% Drop-in for src/primal/ManiSDP.m:6 (SeDuMi data, arbitrary affine constraints) on the B200 engine.
if nargin < 5; options = struct(); end
[X, obj, data] = manisdp_b200_driver(3, [], At, b, c, K, options);
end
