cd /root/repo
timeout 900 python -m pytest tests/test_gpu_facade.py -m gpu -x -q -s -k "complex" 2>&1 | tail -25
