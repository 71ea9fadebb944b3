timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --set full --import-source on --clock-control none -k regex:b2_rti_solve_wpp -s 1 -c 1 -f -o gpurun_out/prof_wpp_final python scripts/prof_solve.py 4096 2 2 > gpurun_out/prof_final.log 2>&1; tail -1 gpurun_out/prof_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b.log 2>&1
python scripts/bench_configs.py > gpurun_out/secondary.json 2> gpurun_out/secondary.err
