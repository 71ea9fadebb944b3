set -x
timeout 900 ncu --set full --import-source on --clock-control none -k regex:b2_rti_solve_wpp -s 1 -c 1 -f -o gpurun_out/r02b_wpp python scripts/prof_solve.py 4096 2 2 > gpurun_out/r02b_wpp.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:b2_rti_solve_cta -s 1 -c 1 -f -o gpurun_out/r02b_cta python scripts/prof_solve.py 148 4 2 > gpurun_out/r02b_cta.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02b_bench_under_ncu.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed.avg.per_cycle_active,smsp__inst_executed.sum"
timeout 900 ncu --metrics $M --clock-control none -k regex:b2_rti_solve_wpp -s 1 -c 1 --csv --log-file gpurun_out/r02b_n100_metrics.csv python scripts/time_cfg4.py 1024 0 > /dev/null 2>&1
ls -la gpurun_out | tail -8
