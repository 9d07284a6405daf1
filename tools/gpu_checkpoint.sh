#!/bin/bash
# One GPU-box visit: parity tests, per-kernel timings, and one ncu --set full capture of each role of the fused kernel.
#   tools/gpu_checkpoint.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-ckpt}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -4 $out/${tag}_tests.log
for e in 0 1; do python tools/quick_bench.py --nerr $e --tag "$tag e$e"; done
python tools/quick_bench.py --batch 1024 --tag "$tag b1024"
python tools/quick_bench.py --model full_blockaded --tag "$tag d7"
python tools/quick_bench.py --model full_blockaded --nerr 1 --tag "$tag d7 e1"
ncu --set full --import-source on --clock-control none -k regex:k_fused_q -s 3 -c 1 -f -o $out/${tag}_fq_e0 python tools/quick_bench.py --reps 1 > $out/${tag}_ncu_e0.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_fused_q -s 7 -c 1 -f -o $out/${tag}_fq_err python tools/quick_bench.py --reps 1 --nerr 1 > $out/${tag}_ncu_err.log 2>&1
ls -la $out/${tag}_*
