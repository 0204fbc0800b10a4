cd /root/repo
timeout 900 python bench.py --no-variants --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; python - <<'PY'
import json
j=json.loads(open('/root/repo/gpurun_out/bench_now.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","ms_per_step")}, "pcg", j["pcg"]["ms"], j["pcg"]["iterations"], "e2e", j["e2e"], "frac", j["roofline"]["frac"], "torque-only", j["sweep_torque_only"]["s_per_position"])
PY
tail -3 gpurun_out/bench_now.err
