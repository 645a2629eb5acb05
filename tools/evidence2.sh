#!/bin/bash
# final r02 evidence refresh (after the CTA-pair kernels): tests, kernel zoo, launch lists, per-launch GEMM metrics, bench lines
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second"
python tools/kernel_zoo.py > gpurun_out/zoo_plain.json 2> gpurun_out/zoo.err && \
  ncu --set full --clock-control none --profile-from-start off -k "regex:qsample|stage_kernel|timestep_embed|haar|ddpm_step|gn_|layernorm|gemm_|flash_attn|attn_norms|mse_finalize|eps_mse" -f -o /tmp/r02_zoo \
      python tools/kernel_zoo.py > gpurun_out/zoo.json 2>> gpurun_out/zoo.err
ncu -i /tmp/r02_zoo.ncu-rep --page raw --csv > gpurun_out/r02_zoo_raw.csv 2>/dev/null; echo "zoo rc=$? $(wc -c < gpurun_out/r02_zoo_raw.csv)"
for w in unet128 dit; do
  python tools/one_pass.py $w 1 > gpurun_out/pass_$w.json 2> gpurun_out/pass_$w.err && \
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$w.csv \
        python tools/one_pass.py $w 1 > gpurun_out/pass_ncu_$w.json 2>> gpurun_out/pass_$w.err
  echo "launches $w rc=$? $(cat gpurun_out/pass_$w.json)"
done
python tools/one_pass.py unet128 1 > /dev/null 2>&1 && \
  ncu --metrics $M --clock-control none --profile-from-start off -k regex:gemm_tc --csv --log-file gpurun_out/gemm_metrics_unet128.csv \
      python tools/one_pass.py unet128 1 > /dev/null 2>> gpurun_out/pass_unet128.err
echo "gemm metrics rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_unet128_final.json 2> gpurun_out/bench_unet128.err; echo "unet128 rc=$? $(cut -c1-90 gpurun_out/r02_bench_unet128_final.json)"
for w in cifar dit; do
  timeout 400 python bench.py --workload $w --steps 3 --no-cpu > gpurun_out/r02_bench_${w}_1gpu_v2.json 2> gpurun_out/bench_$w.err; echo "$w rc=$? $(cut -c1-90 gpurun_out/r02_bench_${w}_1gpu_v2.json)"
done
