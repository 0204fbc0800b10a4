set -x
mkdir -p gpurun_out
export XFEMM_BENCH_DIR=/tmp/xfemm_bench
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precond ${1:-jacobi}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${1:-jacobi}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/plain.log | cut -c1-600
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 3 -c 1 -o gpurun_out/prof_${1:-jacobi} $CMD > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
