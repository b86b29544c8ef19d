# giant tier: parity tests, dense-config timings, regression bench
mkdir -p gpurun_out
python -m pacybara_b200.build > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02j_pytest.log
for s in 0.03 0.1 0.3 1.0; do
  timeout 600 python tests/scale_check.py dense $s > gpurun_out/r02j_dense_$s.json 2> gpurun_out/r02j_dense_$s.err; echo "dense $s rc=$?"; cat gpurun_out/r02j_dense_$s.json; tail -3 gpurun_out/r02j_dense_$s.err
done
for c in cfg2 cfg4 cfg5; do
  PCB_BENCH_CONFIG=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r02j_bench_$c.json 2> gpurun_out/r02j_$c.err; echo "$c rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/r02j_bench_$c.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'],d['stage_ms'],d['parity']['equal'],d['gpu_launches'])"
done
shard_profile 8:
timeout 300 python tests/shard_profile.py 8 2>&1 | tail -2
