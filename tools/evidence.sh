#!/bin/bash
# one-GPU evidence run for profiles/ (r02): kernel zoo (--set full of every kernel), launch lists, per-launch GEMM metrics,
# kernel-time sums at the strong-scaling launch size, bench lines of the other configs.   gpurun -- 'bash tools/evidence.sh'
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second"
python tools/kernel_zoo.py > gpurun_out/zoo_plain.json 2> gpurun_out/zoo.err && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dcb -f -o gpurun_out/r02_zoo \
      python tools/kernel_zoo.py > gpurun_out/zoo.json 2>> gpurun_out/zoo.err
echo "zoo rc=$?"
for w in unet128 dit cifar; do
  img=1; [ $w = cifar ] && img=16
  python tools/one_pass.py $w $img > gpurun_out/pass_$w.json 2> gpurun_out/pass_$w.err && \
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$w.csv \
        python tools/one_pass.py $w $img > gpurun_out/pass_ncu_$w.json 2>> gpurun_out/pass_$w.err
  echo "launches $w rc=$? $(cat gpurun_out/pass_$w.json)"
done
python tools/one_pass.py unet128 1 > /dev/null 2>&1 && \
  ncu --metrics $M --clock-control none --profile-from-start off -k regex:gemm_tc --csv --log-file gpurun_out/gemm_metrics_unet128.csv \
      python tools/one_pass.py unet128 1 > /dev/null 2>> gpurun_out/pass_unet128.err
echo "gemm metrics rc=$?"
python tools/one_pass.py unet128 1 100 > gpurun_out/pass_unet128_mb100.json 2>/dev/null && \
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet128_mb100.csv \
      python tools/one_pass.py unet128 1 100 > /dev/null 2>&1
echo "mb100 rc=$? $(cat gpurun_out/pass_unet128_mb100.json)"
for w in cifar dit; do
  timeout 400 python bench.py --workload $w --steps 3 --no-cpu > gpurun_out/r02_bench_${w}_1gpu.json 2> gpurun_out/bench_$w.err; echo "$w rc=$? $(cut -c1-100 gpurun_out/r02_bench_${w}_1gpu.json)"
done
