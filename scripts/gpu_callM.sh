# giant-tier knobs: bucket-count rule, warps per SM (via cap), on the dense library and cfg4
mkdir -p gpurun_out
for hb in 0 16 32 64; do
PCB_HUGE_BUCKETS=$hb timeout 300 python tests/scale_check.py dense 0.1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('huge_buckets $hb dense0.1 replay', d['stage_ms']['replay'], 'rounds', d['huge_rounds'], 'comps', d['huge_components'])"
PCB_HUGE_BUCKETS=$hb timeout 300 python tests/scale_check.py dense 0.03 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('huge_buckets $hb dense0.03 replay', d['stage_ms']['replay'], 'rounds', d['huge_rounds'], 'comps', d['huge_components'])"
PCB_HUGE_BUCKETS=$hb PCB_BENCH_CONFIG=cfg4 timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-files-e2e --no-parity 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('huge_buckets $hb cfg4', round(d['ms_per_step'],3), 'replay', round(d['stage_ms']['replay'],3))"
PCB_HUGE_BUCKETS=$hb PCB_BENCH_CONFIG=cfg2 timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-files-e2e --no-parity 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('huge_buckets $hb cfg2', round(d['ms_per_step'],3), 'replay', round(d['stage_ms']['replay'],3))"
done
