# short-barcode index: parity + timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p_pytest.log
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  PCB_BENCH_CONFIG=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02p_bench_$c.json 2> gpurun_out/r02p_$c.err; echo "$c rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02p_bench_$c.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['stage_ms'],d['parity']['equal'],d['gpu_launches'])"
done
