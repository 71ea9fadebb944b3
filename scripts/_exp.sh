for cfg in "12 100" "11 92" "10 84" "9 75" "8 67" "6 50"; do set -- $cfg
B2MPC_EXPERIMENT_OCC=$1 B2MPC_EXPERIMENT_CARVEOUT=$2 python scripts/time_solve.py 32768 2 4 | sed "s/^/occ=$1 carve=$2 /"
done
for cfg in "10 84" "8 67"; do set -- $cfg
B2MPC_EXPERIMENT_OCC=$1 B2MPC_EXPERIMENT_CARVEOUT=$2 python scripts/time_solve.py 4096 2 7 | sed "s/^/occ=$1 carve=$2 /"
done
