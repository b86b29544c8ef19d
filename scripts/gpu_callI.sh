# final-code captures: launch lists (cfg2, cfg4), --set full of the replay + probe kernels at cfg2
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02q_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_q1.log 2>&1; echo "ncu cfg2 rc=$?"
PCB_BENCH_CONFIG=cfg4 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02q_launches_cfg4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_q2.log 2>&1; echo "ncu cfg4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:component_kernel -c 6 -f -o gpurun_out/r02q_replay python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_q3.log 2>&1; echo "ncu replay rc=$?"
ncu --set full --clock-control none --import-source on -k regex:probe_kernel -c 2 -f -o gpurun_out/r02q_probe python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_q4.log 2>&1; echo "ncu probe rc=$?"
