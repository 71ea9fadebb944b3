timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python scripts/time_solve.py 4096 2 9
python scripts/time_solve.py 32768 2 5
python scripts/dump_parity.py 2>&1 | grep "kernel 2"
