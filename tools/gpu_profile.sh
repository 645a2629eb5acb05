mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --images 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 9 -c 3 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | head -20
tail -n 3 gpurun_out/ncu_launches.log gpurun_out/ncu_full.log 2>/dev/null
