cd /root/repo
timeout 900 python scratch/e2e_prof.py 2>&1 | grep -v "^$" | grep "linprob_create\|attachDevice\|host tables\|mag_create\|newton loop\|rep " | tail -12
timeout 1200 python -m pytest tests/test_gpu_static.py tests/test_gpu_linprob.py tests/test_gpu_facade.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
