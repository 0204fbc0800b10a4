cd /root/repo
for o in 3 2; do
XFEMM_EXTRAPOLATION_ORDER=$o timeout 900 python bench.py --no-variants --no-cpu-baseline --steps 12 --warmup 6 > gpurun_out/bench_o$o.json 2> gpurun_out/bench_o$o.err; python - <<PY
import json
j=json.loads(open('/root/repo/gpurun_out/bench_o$o.json').read().strip().splitlines()[-1])
print("order $o", {k:j[k] for k in ("value",)}, "pcg", j["pcg"]["ms"], j["pcg"]["iterations"], "torque-only", j["sweep_torque_only"]["s_per_position"])
PY
done
