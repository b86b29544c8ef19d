# experiments: warps per block of the small kernel; seed table cap (cfg3)
mkdir -p gpurun_out
for w in 4 2 1; do for c in cfg2 cfg4 cfg5; do
PCB_SMALL_WARPS=$w PCB_BENCH_CONFIG=$c timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-files-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('small_warps $w $c', round(d['ms_per_step'],3), 'replay', round(d['stage_ms']['replay'],3), d['parity']['equal'])"
done; done
for b in 26 24 22; do
PCB_SEED_BINS_LOG2=$b PCB_BENCH_CONFIG=cfg3 timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-files-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('seed_bins $b cfg3', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()}, d['run']['seed_len'], d['parity']['equal'])"
done
