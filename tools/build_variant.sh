#!/bin/bash
# Build a variant of the library with extra -D flags applied to selected translation units (kernel A/B experiments).
#   tools/build_variant.sh <name> "<flags>" tu1.cu [tu2.cu ...]   ->  robustgrape_b200/lib/variants/lib<name>.so
set -e
name=$1; flags=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/robustgrape_b200/lib/variants; mkdir -p $out/obj_$name
objs=""
for src in $root/robustgrape_b200/csrc/*.cu; do
  b=$(basename $src .cu); use=$root/robustgrape_b200/lib/obj/$b.o
  for tu in "$@"; do
    if [ "$tu" == "$b.cu" ]; then
      use=$out/obj_$name/$b.o
      nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -diag-suppress 68,128,20058 -Xcompiler -fPIC $flags -c -o $use $src &
    fi
  done
  objs="$objs $use"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $out/lib$name.so $objs
echo built $out/lib$name.so
