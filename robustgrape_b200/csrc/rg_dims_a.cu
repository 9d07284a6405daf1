// Instantiates the kernels and launch templates for ndim = 2, 3, 4 (see rg_host.cuh).
#include "rg_host.cuh"
RG_DEFINE_DIM(2)
RG_DEFINE_DIM(3)
RG_DEFINE_DIM(4)
