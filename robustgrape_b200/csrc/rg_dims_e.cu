// Instantiates the kernels and launch templates for ndim = 10 (see rg_host.cuh).
#include "rg_host.cuh"
RG_DEFINE_DIM(10)
