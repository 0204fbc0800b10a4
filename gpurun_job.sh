cd /root/repo
timeout 900 python -m pytest tests/test_gpu_static.py tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -3
XFEMM_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-variants >gpurun_out/v.log 2>gpurun_out/timing.log; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/v.log').read().strip().splitlines()[-1])
print("value", round(d['value']*1e3,2), "e2e", round(d['e2e']['value'],3), "amg pcg ms", round(d['pcg']['ms'],2), "its", d['pcg']['iterations'], "frac", round(d['roofline']['frac'],3))
PY
grep -E "total|pattern|newton" gpurun_out/timing.log | head -6
