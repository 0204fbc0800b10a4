cd /root/repo
timeout 1200 python scratch/nl_big.py 0.00925 2>&1 | tail -3
