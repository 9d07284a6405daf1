// three-kernel block-2 path, pattern 2 (see rg_block2_patterns.cuh)
#define RG_B2_D 5
#define RG_B2_ID 2
#define RG_B2_MASK B2_M5_FULL
#include "rg_b2_impl.inl"
