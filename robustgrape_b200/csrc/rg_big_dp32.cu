// dense path (rg_big.cuh) for ndim padded to 32
#define RG_BIG_DP 32
#include "rg_big_impl.inl"
