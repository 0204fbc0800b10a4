cd /root/repo
timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-variants --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; python - <<'PY'
import json
j=json.loads(open('/root/repo/gpurun_out/bench_now.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","ms_per_step")}, "pcg", j["pcg"]["ms"], j["pcg"]["iterations"], "asm", j["assembly"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "torque-only", j["sweep_torque_only"]["s_per_position"], j["sweep_torque_only"]["cold_start"])
PY
tail -3 gpurun_out/bench_now.err
