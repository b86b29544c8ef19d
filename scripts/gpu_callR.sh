# cfg5 (2*10^7 reads) sharded over N GPUs: the strong-scaling point BASELINE.json names
mkdir -p gpurun_out
N=${1:-8}
export PCB_BENCH_CONFIG=cfg5 PCB_BENCH_SCALE=$(python -c "print(1.0/$N)")
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02ac_bench_cfg5_n$N.json 2> gpurun_out/r02ac_bench_cfg5_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02ac_bench_cfg5_n$N.json') if l.startswith('{')][-1])
print(d['n_gpus'],d['config']['workload'],d['value'],d['ms_per_step'],d['e2e'],d['stage_ms'],d.get('parity',{}).get('equal'),d.get('shard_balance',{}).get('pairs_scored'),d.get('stage_limit'))
PY
tail -3 gpurun_out/r02ac_bench_cfg5_n$N.err
