cd /root/repo
timeout 900 python -m pytest tests/test_gpu_static.py -m gpu -x -q -k "tbmag" 2>&1 | tail -8
