set -x
cd /root/repo
mkdir -p gpurun_out
export XFEMM_BENCH_DIR=/tmp/xfemm_bench
P=${1:-amg}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --precond $P"
$CMD > gpurun_out/plain_$P.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$P.csv $CMD > gpurun_out/ncu_launch_$P.log 2>&1
tail -1 gpurun_out/plain_$P.log | cut -c1-400
$CMD > gpurun_out/plain2_$P.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 2 -c 1 -f -o gpurun_out/prof_$P $CMD > gpurun_out/ncu_full_$P.log 2>&1
tail -3 gpurun_out/ncu_full_$P.log
ls -la gpurun_out | tail -8
