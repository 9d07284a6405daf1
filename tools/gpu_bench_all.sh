#!/bin/bash
# Round-end 1-GPU measurement set: contract bench, reference arm, the other workloads, ncu launch list of the bench command.
tag=${1:-r02}
out=gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 600 $out/${tag}_bench_1gpu.json
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>/dev/null
python bench.py --workload C5 --steps 2 --warmup 3 > $out/${tag}_bench_C5_d64.json 2> $out/${tag}_C5.err
python bench.py --workload d16 --steps 2 --warmup 3 > $out/${tag}_bench_d16.json 2> $out/${tag}_d16.err
python bench.py --workload C3 --steps 20 --warmup 3 > $out/${tag}_bench_C3.json 2> $out/${tag}_C3.err
python bench.py --nerr 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $out/${tag}_bench_C4prime_e1.json 2> $out/${tag}_e1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
ls -la $out/${tag}_*
