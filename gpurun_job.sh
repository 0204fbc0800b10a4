cd /root/repo
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -s 2>&1 | grep -E "iterations|passed|failed|Error|error" | cut -c1-400 | tail -12
for P in amg jacobi; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --partitioned --precond $P > gpurun_out/part2_$P.log 2>&1; tail -1 gpurun_out/part2_$P.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['preconditioner'][:20], 'value', d['value'], d['pcg'])"; done
XFEMM_DIST_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --partitioned --precond amg > gpurun_out/part2_graph.log 2>&1; tail -2 gpurun_out/part2_graph.log | cut -c1-600
