cd /root/repo
( time python bench.py ) > gpurun_out/bench_default.log 2>&1; grep "^real" gpurun_out/bench_default.log; grep "^{" gpurun_out/bench_default.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','steps','warmup','e2e','gpu_launches','clocks')}); print(d['pcg']['ms'], d['pcg']['iterations']); print(d['roofline']['frac'], d['roofline']['traffic']); print(d['variant']); print(d['sweep_torque_only']['s_per_position'], d['sweep_torque_only']['cold_start'])"
bash gpurun_prof.sh amg > /dev/null 2>&1
ls -la gpurun_out/prof_amg.ncu-rep
