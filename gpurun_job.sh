cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
XFEMM_TIMING=1 timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/timing.log >gpurun_out/bench_amg.log; tail -1 gpurun_out/bench_amg.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','e2e','roofline')})"; grep -E "release|total|pattern|AMG|SELL|create|attach|mag_" gpurun_out/timing.log | tail -22
