cd /root/repo
timeout 900 python -m pytest tests/test_gpu_harmonic.py tests/test_gpu_facade.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python scratch/newton_big.py 0.02 2>&1 | tail -4
