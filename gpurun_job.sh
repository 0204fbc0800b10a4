cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/sweep2.log 2>&1; tail -1 gpurun_out/sweep2.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','e2e','host_setup_s','clocks')}); print(d['pcg'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --partitioned > gpurun_out/part2.log 2>&1; tail -1 gpurun_out/part2.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -2
