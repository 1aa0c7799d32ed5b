function [X, obj, data] = ManiSDP_onlyunitdiag(C, options)
% Drop-in for src/primal/ManiSDP_onlyunitdiag.m:6 -- min <C,X> s.t. diag(X) = 1, X >= 0 -- on the B200 engine.
if nargin < 2; options = struct(); end
[X, obj, data] = manisdp_b200_driver(0, C, [], [], [], [], options);
end
