mkdir -p gpurun_out
timeout 1200 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -n 15 gpurun_out/bench.err
