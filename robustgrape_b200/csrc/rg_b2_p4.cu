// three-kernel block-2 path, pattern 4 (see rg_block2_patterns.cuh)
#define RG_B2_D 7
#define RG_B2_ID 4
#define RG_B2_MASK B2_M7_FULL
#include "rg_b2_impl.inl"
