#!/bin/bash
# 8-GPU box: multi-GPU parity test, strong-scaling curve of unet-128 (one batch of 4 images split over N ranks), and the
# configs the north star names for 8 GPUs (unet-256, IPMSA DWT) at N = 1 and N = 8.   gpurun --gpus 8 -- 'bash tools/scale_all.sh r02'
tag=${1:-rNN}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_g_multigpu.py -m gpu -q 2>&1 | tail -3
run() {  # n, out, args...
  n=$1; out=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/$out 2> gpurun_out/$out.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n "$@" > gpurun_out/$out 2> gpurun_out/$out.err
  fi
  echo "$out rc=$? $(cut -c1-120 gpurun_out/$out)"
}
for n in 1 2 4 8; do run $n ${tag}_bench_unet128_strong_${n}gpu.json --steps 10 --warmup 3 --no-cpu --scaling strong; done
for n in 1 2 4 8; do run $n ${tag}_bench_unet128_strong16_${n}gpu.json --steps 5 --warmup 3 --no-cpu --scaling strong --images 16; done
for w in unet256 ipmsa; do
  run 1 ${tag}_bench_${w}_1gpu.json --workload $w --steps 3 --warmup 3 --no-cpu
  run 8 ${tag}_bench_${w}_8gpu.json --workload $w --steps 3 --warmup 3 --no-cpu
done
run 8 ${tag}_bench_unet128_weak_8gpu.json --steps 10 --warmup 3 --no-cpu
cat gpurun_out/r02_multigpu_check_*.json
