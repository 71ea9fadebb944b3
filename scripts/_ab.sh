L=dynamic-obstacle-avoidance-mpc_b200/lib
for rep in 1 2; do
for v in "" _prev; do
  B2MPC_LIB=$L/libb200mpc$v.so python scripts/time_solve.py 4096 2 11 2>&1 | tail -1 | cut -c1-40,100-170 | sed "s/^/[$v] /"
  B2MPC_LIB=$L/libb200mpc$v.so python scripts/time_solve.py 32768 2 5 2>&1 | tail -1 | cut -c1-40,100-170 | sed "s/^/[$v] /"
done; done
for v in "" _prev; do B2MPC_LIB=$L/libb200mpc$v.so python scripts/time_cfg4.py 4096 0 2>&1 | tail -1; done
