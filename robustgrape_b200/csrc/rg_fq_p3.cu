// one-launch fused quaternion path, pattern 3 (see rg_block2_patterns.cuh)
#define RG_B2_D 7
#define RG_B2_ID 3
#define RG_B2_MASK B2_M7_DRIVE
#include "rg_fq_impl.inl"
