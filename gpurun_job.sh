cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/v.log 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/v.log').read().strip().splitlines()[-1])
print("value", round(d['value']*1e3,2), "e2e", round(d['e2e']['value'],3), "pcg ms", round(d['pcg']['ms'],2), "its", d['pcg']['iterations'], "frac", round(d['roofline']['frac'],3), [(v['preconditioner'], round(v['pcg_ms'],1), v['iterations'], round(v['algorithmic_gbs'])) for v in d['variant']], d['sweep_torque_only']['s_per_position'], d['sweep_torque_only']['cold_start'])
PY
