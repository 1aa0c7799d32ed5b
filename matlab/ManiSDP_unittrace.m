function [X, obj, data] = ManiSDP_unittrace(At, b, c, K, options)
% Drop-in for src/primal/ManiSDP_unittrace.m:7 (SeDuMi data, unit trace + affine constraints) on the B200 engine.
if nargin < 5; options = struct(); end
[X, obj, data] = manisdp_b200_driver(2, [], At, b, c, K, options);
end
