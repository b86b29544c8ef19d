# final single-GPU campaign: tests, driver-style lines, all configs, captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
for c in cfg1 cfg3 cfg4 cfg5; do
  PCB_BENCH_CONFIG=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02u_bench_$c.json 2> gpurun_out/r02u_$c.err; echo "$c rc=$?"
done
for s in 0.3 1.0; do timeout 600 python tests/scale_check.py dense $s > gpurun_out/r02u_dense_$s.json 2>/dev/null; echo "dense $s rc=$?"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02u_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_u1.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:component_kernel -c 6 -f -o gpurun_out/r02u_replay python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_u2.log 2>&1; echo "ncu replay rc=$?"
ncu --set full --clock-control none --import-source on -k regex:probe_kernel -c 2 -f -o gpurun_out/r02u_probe python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-files-e2e > gpurun_out/ncu_u3.log 2>&1; echo "ncu probe rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02u_bench.json') if l.startswith('{')][-1])
print('cfg2', d['ms_per_step'], d['e2e'], d['stage_ms'], d['files_e2e'], d['cpu_baseline']['value'], d['roofline']['frac'], d['parity']['equal'])
for c in ('cfg1','cfg3','cfg4','cfg5'):
    d=json.load(open(f'gpurun_out/r02u_bench_{c}.json')); print(c, round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()}, d['parity']['equal'])
for s in ('0.3','1.0'):
    d=json.load(open(f'gpurun_out/r02u_dense_{s}.json')); print('dense',s, d['wall_ms'], d['stage_ms'], d['probe_ms'], d['huge_rounds'])
PY
