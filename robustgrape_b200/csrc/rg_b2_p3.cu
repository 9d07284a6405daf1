// three-kernel block-2 path, pattern 3 (see rg_block2_patterns.cuh)
#define RG_B2_D 7
#define RG_B2_ID 3
#define RG_B2_MASK B2_M7_DRIVE
#include "rg_b2_impl.inl"
