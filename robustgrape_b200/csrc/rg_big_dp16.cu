// dense path (rg_big.cuh) for ndim padded to 16
#define RG_BIG_DP 16
#include "rg_big_impl.inl"
