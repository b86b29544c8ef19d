set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02i_bench_cfg2.json 2> gpurun_out/r02i_cfg2.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02i_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_a.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:component_kernel -c 6 -f -o gpurun_out/r02i_replay python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_b.log 2>&1; echo "ncu replay rc=$?"
ncu --set full --clock-control none --import-source on -k regex:probe_kernel -c 2 -f -o gpurun_out/r02i_probe python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_c.log 2>&1; echo "ncu probe rc=$?"
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
