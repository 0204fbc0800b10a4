cd /root/repo
timeout 900 python -m pytest tests/test_gpu_harmonic.py -m gpu -x -q -s 2>&1 | grep -i "iterations\|passed\|failed\|error" | tail -8
HC_FIRST=2 timeout 1500 python scratch/harm_cases.py 2>&1 | tail -30
