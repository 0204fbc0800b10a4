cd /root/repo
timeout 900 python -m pytest tests/test_gpu_harmonic.py tests/test_gpu_facade.py tests/test_gpu_cnewton.py -m gpu -x -q 2>&1 | tail -4
HC_FIRST=2 timeout 1500 python scratch/harm_cases.py 2>&1 | grep "precond 2\|rel diff" | tail -12
