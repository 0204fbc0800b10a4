cd /root/repo
timeout 300 python -m pytest tests/test_gpu_linprob.py tests/test_gpu_static.py -m gpu -x -q 2>&1 | tail -3
for R in 64 1024; do XFEMM_AMG_LOCAL_ROWS=$R XFEMM_AMG_TRACE=1 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --precond amg 2>gpurun_out/trace.log >gpurun_out/sweep_amg.log; grep -o '"pcg": {[^}]*}' gpurun_out/sweep_amg.log | tail -1; tail -1 gpurun_out/trace.log; done
