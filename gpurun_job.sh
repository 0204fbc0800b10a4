cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline >gpurun_out/v.log 2>&1; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/v.log').read().strip().splitlines()[-1])
print("value", round(d['value']*1e3,2), "e2e", round(d['e2e']['value'],3), "amg pcg ms", round(d['pcg']['ms'],2), "its", d['pcg']['iterations'], "frac", round(d['roofline']['frac'],3), "traffic", d['roofline']['traffic'], "|", [ (v['preconditioner'], round(v['pcg_ms'],1), round(v['ms_per_iter'],4), round(v['algorithmic_gbs'])) for v in d['variant']], d['sweep_torque_only']['s_per_position'], d['clocks'])
PY
