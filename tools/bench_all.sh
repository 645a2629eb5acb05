#!/bin/bash
# every bench line committed under profiles/ for a round: the five BASELINE configs, the SURVEY 8(f) paths and the CPU arm.
#   gpurun -- 'bash tools/bench_all.sh v7'
tag=${1:-vN}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r01_bench_unet128_$tag.json 2> gpurun_out/bench_unet128.err; echo "unet128 rc=$?"
for w in cifar unet256 dit ipmsa; do
  timeout 400 python bench.py --workload $w --steps 3 --no-cpu > gpurun_out/r01_bench_${w}_$tag.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
for p in sample loss evaluate; do
  timeout 300 python bench.py --path $p --steps 3 > gpurun_out/r01_bench_path_${p}_$tag.json 2> gpurun_out/bench_path_$p.err; echo "$p rc=$?"
done
timeout 300 python bench.py --path sample --steps 3 --images 32 > gpurun_out/r01_bench_path_sample32_$tag.json 2>/dev/null
timeout 600 python bench.py --workload cifar --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_cifar_reference_arm_$tag.json 2>/dev/null
for f in gpurun_out/r01_bench_*_$tag.json; do echo "$f: $(cut -c1-95 $f)"; done
