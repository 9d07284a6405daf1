// dense path (rg_big.cuh) for ndim padded to 48
#define RG_BIG_DP 48
#include "rg_big_impl.inl"
