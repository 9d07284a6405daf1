// Instantiates the kernels and launch templates for ndim = 6, 7 (see rg_host.cuh).
#include "rg_host.cuh"
RG_DEFINE_DIM(6)
RG_DEFINE_DIM(7)
