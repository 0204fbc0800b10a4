cd /root/repo
timeout 900 python -m pytest tests/test_gpu_cnewton.py -m gpu -x -q -s 2>&1 | tail -25
