set -x
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --precond mcssor > gpurun_out/bench_mcssor.json 2> gpurun_out/bench_mcssor.err; tail -3 gpurun_out/bench_mcssor.err; cat gpurun_out/bench_mcssor.json
python bench.py --steps 3 --warmup 3 --precond jacobi --no-cpu-baseline > gpurun_out/bench_jacobi.json 2> gpurun_out/bench_jacobi.err; tail -3 gpurun_out/bench_jacobi.err; cat gpurun_out/bench_jacobi.json
