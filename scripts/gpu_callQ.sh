# how the small replay kernel's time depends on its occupancy (extra shared memory per block lowers it)
for pad in 0 2048 8192 16384; do
PCB_SMALL_PAD=$pad timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-files-e2e --no-parity 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('small_pad $pad cfg2', round(d['ms_per_step'],3), 'replay', round(d['stage_ms']['replay'],3))"
PCB_SMALL_PAD=$pad PCB_BENCH_CONFIG=cfg5 timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-files-e2e --no-parity 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('small_pad $pad cfg5', round(d['ms_per_step'],3), 'replay', round(d['stage_ms']['replay'],3))"
done
