# class sub-bins in the probe: parity + timing; cfg5 launch list; N=8 shard profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_match.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest.log
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "cfg5 or cfg4" > gpurun_out/r02m_pytest2.log 2>&1; echo "pytest scale rc=$?"; tail -3 gpurun_out/r02m_pytest2.log
for c in cfg2 cfg3 cfg4 cfg5; do
  PCB_BENCH_CONFIG=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02m_bench_$c.json 2> gpurun_out/r02m_$c.err; echo "$c rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02m_bench_$c.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['stage_ms'],d['parity']['equal'],d['gpu_launches'])"
done
PCB_PROBE_CLASSES=0 PCB_BENCH_CONFIG=cfg5 timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-files-e2e --no-parity 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 classes off', d['ms_per_step'], d['stage_ms'])"
PCB_BENCH_CONFIG=cfg5 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02m_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_m.log 2>&1; echo "ncu rc=$?"
timeout 300 python tests/shard_profile.py 8 2>&1 | tail -2
