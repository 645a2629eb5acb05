#!/bin/bash
# ncu evidence for profiles/ (B200_PROFILING.md recipe): run under gpurun, one GPU.
#   tools/profile.sh launches   -> gpurun_out/launches.csv  (every launch of one classify pass with its device time)
#   tools/profile.sh full       -> gpurun_out/prof_gemm.ncu-rep (--set full of 3 gemm_tc2 launches)
#   tools/profile.sh gn         -> gpurun_out/prof_gn.ncu-rep   (--set full of 3 gn_apply launches)
set -e
CMD="python bench.py --images 1 --steps 1 --warmup 3 --no-cpu"
export DCB_CUDA_GRAPH=0
case "$1" in
  launches)
    $CMD > gpurun_out/plain.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1 ;;
  full)
    $CMD > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 198 -c 6 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1 ;;
  gn)
    $CMD > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:gn_apply_kernel -s 150 -c 8 -f -o gpurun_out/prof_gn $CMD > gpurun_out/ncu_gn.log 2>&1 ;;
esac
# tools/profile.sh traffic -> gpurun_out/gemm_traffic.csv: DRAM bytes, tensor-pipe activity, L2 hit rate of every tcgen05 GEMM launch of one pass
if [ "$1" = "traffic" ]; then
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
      --clock-control none -k regex:gemm_tc -s 369 -c 123 --csv --log-file gpurun_out/gemm_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
fi
# tools/profile.sh attn -> gpurun_out/prof_attn.ncu-rep (--set full of 2 flash_attn_tc launches of the DiT-B/4 workload)
if [ "$1" = "attn" ]; then
  ACMD="python tools/kernel_times.py dit 1"
  $ACMD > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -s 14 -c 2 -f -o gpurun_out/prof_attn $ACMD > gpurun_out/ncu_attn.log 2>&1
fi
# tools/profile.sh domgemm -> gpurun_out/prof_domgemm.ncu-rep: --set full of the dominant conv shape (256 -> 128 ch 3x3 @128^2, x-halo gemm_tc2) and
#                             gpurun_out/prof_attn.ncu-rep: the tcgen05 attention kernel at the DiT-B/4 shape
if [ "$1" = "domgemm" ]; then
  G="python tools/gemm_micro.py conv256to128"; A="python tools/attn_micro.py"
  export MICRO_ITERS=3
  $G > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 4 -c 1 -f -o gpurun_out/prof_domgemm $G > gpurun_out/ncu_domgemm.log 2>&1
  $A > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -s 4 -c 1 -f -o gpurun_out/prof_attn $A > gpurun_out/ncu_attn.log 2>&1
fi
