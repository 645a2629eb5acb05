mkdir -p gpurun_out
rm -f gpurun_out/rc.txt
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_a_kernels.py -m gpu -q -rA --timeout 300 -p no:cacheprovider > gpurun_out/a.log 2>&1; echo "A rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_b_tcgen05.py -m gpu -q -rA --timeout 120 -p no:cacheprovider > gpurun_out/b.log 2>&1; echo "B rc=$?" >> gpurun_out/rc.txt
timeout 1200 python -m pytest tests/test_gpu_c_models.py -m gpu -q -rA --timeout 600 -p no:cacheprovider > gpurun_out/c.log 2>&1; echo "C rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
for f in a b c; do tail -n 4 gpurun_out/$f.log; done
