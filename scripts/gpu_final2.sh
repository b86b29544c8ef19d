# last check of the tree as it stands: GPU suite, smoke, the driver's bench command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ab_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02ab_bench.json') if l.startswith('{')][-1])
print('cfg2', d['steps'], d['ms_per_step'], d['e2e'], d['stage_ms'], d['files_e2e'], d['cpu_baseline']['value'], d['roofline'], d['parity']['equal'], d['gpu_launches'], d['clocks'])
PY
