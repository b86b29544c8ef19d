# cluster drop-in through the device loaders + device writer: cli tests, three-stage timing, files_e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_ingest.py -m gpu -x -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02t_pytest.log
timeout 600 python scripts/time_dropins.py cfg2 1.0 > gpurun_out/r02t_dropins.json 2> gpurun_out/r02t_dropins.err; echo "dropins rc=$?"; cat gpurun_out/r02t_dropins.json; tail -3 gpurun_out/r02t_dropins.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02t_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('files_e2e'))"
