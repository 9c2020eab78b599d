function [v, usediters] = projfunc(s, k1, k2, nn)
% Drop-in for projfunc.m:1 through libnmfb200.so.
[v, usediters] = nmfb_mex('projfunc', single(s(:)), k1, k2, nn);
end
