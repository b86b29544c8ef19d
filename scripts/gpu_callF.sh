# N GPUs: sharded path vs single-GPU arrays and the oracle, bench lines
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py cfg1 1.0 > gpurun_out/r02n_mgcheck_n$N.log 2>&1; echo "multigpu_check rc=$?"; tail -3 gpurun_out/r02n_mgcheck_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02n_bench_n$N.json 2> gpurun_out/r02n_bench_n$N.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r02n_bench_n$N.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e'],d['stage_ms'],d.get('parity'))"
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or fused_cli" 2>&1 | tail -2
