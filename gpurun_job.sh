set -x
cd /root/repo
python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --partitioned > gpurun_out/part2.log 2>&1; tail -3 gpurun_out/part2.log
python bench.py --gpus 1 --steps 3 --warmup 3 --partitioned > gpurun_out/part1.log 2>&1; tail -3 gpurun_out/part1.log
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep1.log 2>&1; tail -1 gpurun_out/sweep1.log
XFEMM_BENCH_NO_UNION=1 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sweep1_nounion.log 2>&1; tail -1 gpurun_out/sweep1_nounion.log
