cd /root/repo
timeout 900 python -m pytest tests/test_gpu_linprob.py -m gpu -x -q -k pattern 2>&1 | tail -8
