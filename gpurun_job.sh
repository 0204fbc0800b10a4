set -x
cd /root/repo
python -m pytest tests/test_gpu_linprob.py tests/test_gpu_static.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep1.log 2>&1; tail -1 gpurun_out/sweep1.log
