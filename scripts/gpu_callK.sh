# device table writer: ingest / cli tests, files_e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ingest.py tests/test_cli.py tests/test_abi.py -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02s_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02s_bench.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02s_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('files_e2e'))"
