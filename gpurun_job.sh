cd /root/repo
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/e2e_prof.py 2>&1 | tail -12
XFEMM_AMG_TRACE=1 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --precond amg 2>gpurun_out/trace.log > gpurun_out/sweep_amg.log; tail -1 gpurun_out/sweep_amg.log; tail -2 gpurun_out/trace.log
