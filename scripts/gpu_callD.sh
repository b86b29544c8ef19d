# pack staging, probe staging experiment (cfg5), shard profile, quick parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_match.py tests/test_abi.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02l_pytest.log
for c in cfg2 cfg5; do
  PCB_BENCH_CONFIG=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02l_bench_$c.json 2> gpurun_out/r02l_$c.err; echo "$c rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02l_bench_$c.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['stage_ms'],d['parity']['equal'],d['gpu_launches'])"
done
PCB_PROBE_STAGING=1 PCB_BENCH_CONFIG=cfg5 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02l_bench_cfg5_staged.json 2> gpurun_out/r02l_cfg5_staged.err; echo "cfg5 staged rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02l_bench_cfg5_staged.json'));print('staged',d['ms_per_step'],d['stage_ms'],d['parity']['equal'])"
PCB_BENCH_CONFIG=cfg5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_kernel -c 1 -f -o gpurun_out/r02l_probe_cfg5_stream python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_l1.log 2>&1; echo "ncu stream rc=$?"
PCB_PROBE_STAGING=1 PCB_BENCH_CONFIG=cfg5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_kernel -c 1 -f -o gpurun_out/r02l_probe_cfg5_staged python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_l2.log 2>&1; echo "ncu staged rc=$?"
timeout 300 python tests/shard_profile.py 8 2>&1 | tail -2
