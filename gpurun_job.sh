cd /root/repo
timeout 900 python -m pytest tests/test_gpu_static.py tests/test_gpu_sweep.py tests/test_gpu_harmonic.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-variants >gpurun_out/bench_amg.log 2>&1; tail -1 gpurun_out/bench_amg.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','e2e','assembly','gpu_launches','sweep_torque_only')}); print(d['pcg']['ms'], d['pcg']['iterations'])"
