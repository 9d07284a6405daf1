// Instantiates the kernels and launch templates for ndim = 8, 9 (see rg_host.cuh).
#include "rg_host.cuh"
RG_DEFINE_DIM(8)
RG_DEFINE_DIM(9)
