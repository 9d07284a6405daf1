for w in 1 2 4; do RG_WPP=$w python tools/quick_bench.py --batch 1024 --tag "b1024 wpp=$w"; done
python tools/quick_bench.py --batch 1024 --tag "b1024 auto"
for w in 1 2 4; do RG_WPP=$w python tools/quick_bench.py --batch 2048 --tag "b2048 wpp=$w"; done
for w in 1 2 4; do RG_WPP=$w python tools/quick_bench.py --batch 4096 --tag "b4096 wpp=$w"; done
python tools/quick_bench.py --batch 4096 --tag "b4096 auto"
ncu --set full --import-source on --clock-control none -k regex:k_fused_q -s 3 -c 1 -f -o gpurun_out/r02c_fq_b1024 python tools/quick_bench.py --reps 1 --batch 1024 > gpurun_out/r02c_ncu_b1024.log 2>&1
