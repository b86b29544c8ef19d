# N GPUs: bench line only
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02w_bench_n$N.json 2> gpurun_out/r02w_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02w_bench_n$N.json') if l.startswith('{')][-1])
print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e'],d['stage_ms'],d.get('parity',{}).get('equal'),d.get('shard_balance'),d.get('stage_limit'))
PY
