set -x
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25
for pc in jacobi mcssor; do
python bench.py --steps 3 --warmup 3 --precond $pc --no-cpu-baseline > gpurun_out/bench_$pc.json 2> gpurun_out/bench_$pc.err; tail -3 gpurun_out/bench_$pc.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$pc.json'));print('$pc',d['value'],d['pcg'],d['roofline']['frac'],d['e2e'])"
done
XFEMM_JACOBI_CLASSIC=1 python bench.py --steps 3 --warmup 3 --precond jacobi --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('classic',d['value'],d['pcg'],d['roofline']['frac'])"
XFEMM_MAX_COLORS=64 python bench.py --steps 3 --warmup 3 --precond mcssor --no-cpu-baseline | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('mcssor-uncapped',d['value'],d['pcg'],d['roofline']['frac'])"
