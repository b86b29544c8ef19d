# pack kernel without the compare chain: parity + bucket-stage timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_match.py tests/test_gpu_ingest.py -m gpu -x -q > gpurun_out/r02ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ad_pytest.log
for c in cfg2 cfg5; do
  PCB_BENCH_CONFIG=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02ad_bench_$c.json 2> gpurun_out/r02ad_$c.err; echo "$c rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02ad_bench_$c.json'));print(round(d['ms_per_step'],3),round(d['e2e']['ms_per_step'],3),{k: round(v,3) for k,v in d['stage_ms'].items()},d['parity']['equal'])"
done
