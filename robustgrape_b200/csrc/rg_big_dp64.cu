// dense path (rg_big.cuh) for ndim padded to 64
#define RG_BIG_DP 64
#include "rg_big_impl.inl"
