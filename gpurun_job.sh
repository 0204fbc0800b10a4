cd /root/repo
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_static.py -m gpu -x -q 2>&1 | tail -3
for O in 2 1; do XFEMM_EXTRAPOLATION_ORDER=$O timeout 600 python bench.py --gpus 1 --steps 8 --warmup 4 --no-cpu-baseline --no-variants >gpurun_out/v.log 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/v.log').read().strip().splitlines()[-1])
print("value", round(d['value']*1e3,2), "amg pcg ms", round(d['pcg']['ms'],2), "its", d['pcg']['iterations'], "relres", d['pcg']['relres'], d['sweep_torque_only']['s_per_position'])
PY
done
