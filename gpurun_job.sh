cd /root/repo
timeout 300 python -m pytest tests/test_gpu_linprob.py tests/test_gpu_static.py -m gpu -x -q 2>&1 | tail -2
free -g | head -2; nproc
( time timeout 1400 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline --no-variants --mesh-size 0.00057 ) > gpurun_out/bench_20M.log 2> gpurun_out/bench_20M.err; tail -1 gpurun_out/bench_20M.log | cut -c1-3000; tail -5 gpurun_out/bench_20M.err
