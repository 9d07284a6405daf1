// Patterns instantiated ahead of time (upper-triangle bit = k(k+1)/2 + i):
//   d = 5 symmetric-blockaded model (src/RydbergTools.jl:31-39): drive (1,3),(2,4) [+ Rydberg diagonal (3,3),(4,4)]
//   d = 7 full-blockaded model (src/RydbergTools.jl:71-81): drive (1,4),(2,5),(3,6) [+ diagonal (4,4),(5,5),(6,6)]
#pragma once
constexpr unsigned B2_M5_DRIVE = (1u << 7) | (1u << 12);
constexpr unsigned B2_M5_FULL = B2_M5_DRIVE | (1u << 9) | (1u << 14);
constexpr unsigned B2_M7_DRIVE = (1u << 11) | (1u << 17) | (1u << 24);
constexpr unsigned B2_M7_FULL = B2_M7_DRIVE | (1u << 14) | (1u << 20) | (1u << 27);
// pattern ids returned by rg_b2_pattern(): 1 = M5_DRIVE, 2 = M5_FULL, 3 = M7_DRIVE, 4 = M7_FULL
struct B2Ops {
    int (*agg)(rg_problem*, const DevProblem&, int, int, int, const double*);
    int (*grad)(rg_problem*, const DevProblem&, int, int, int, const double*, double*, double);
    int (*grad_err)(rg_problem*, const DevProblem&, int, int, int, const double*, double*);
    void (*occupancy)(const rg_problem*, int*, int*);
};
struct FQOps {
    int (*launch)(rg_problem*, const DevProblem&, int, const double*, int, double*, int, double*, double, double, int, const PeerOut*, const FQAccum*);
    int (*prepare)(rg_problem*);
};
extern const B2Ops rg_b2_ops_p1, rg_b2_ops_p2, rg_b2_ops_p3, rg_b2_ops_p4;
extern const FQOps rg_fq_ops_p1, rg_fq_ops_p3;
