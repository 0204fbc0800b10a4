cd /root/repo
timeout 900 python -m pytest tests/test_gpu_harmonic.py -m gpu -x -q -s -k "newton" 2>&1 | tail -25
