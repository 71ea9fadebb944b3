for i in 1 2; do
python scripts/time_solve.py 4096 2 11
B2MPC_EXPERIMENT_ORDER_MODE=1 python scripts/time_solve.py 4096 2 11
done
python scripts/time_solve.py 4096 2 11 7
B2MPC_EXPERIMENT_ORDER_MODE=1 python scripts/time_solve.py 4096 2 11 7
