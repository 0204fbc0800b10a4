cd /root/repo
timeout 900 python -m pytest tests/test_gpu_harmonic.py -m gpu -x -q -s 2>&1 | tail -5
timeout 600 python scratch/harm_big.py 0.0018 2>&1 | tail -1
HARM_PRECOND=2 timeout 600 python scratch/harm_big.py 0.0018 2>&1 | tail -1
