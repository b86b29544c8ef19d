# giant tier: where the time goes (dense 0.3), knob sweep
mkdir -p gpurun_out
PCB_TRACE=1 timeout 300 python tests/scale_check.py dense 0.3 > gpurun_out/r02k_dense_0.3.json 2> gpurun_out/r02k_dense_0.3.trace; cat gpurun_out/r02k_dense_0.3.json; grep "giant" gpurun_out/r02k_dense_0.3.trace | tail -3
for w in 4096 131072; do PCB_HUGE_WINDOW=$w timeout 300 python tests/scale_check.py dense 0.3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('window $w', d['stage_ms']['replay'], d['huge_rounds'])"; done
for c in 256 8192; do PCB_HUGE_CAP=$c timeout 300 python tests/scale_check.py dense 0.3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('cap $c', d['stage_ms']['replay'], d['huge_rounds'])"; done
for r in 64 1024 100000000; do PCB_HUGE_READS=$r timeout 300 python tests/scale_check.py dense 0.1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('reads-threshold $r (dense 0.1)', d['stage_ms']['replay'], d['huge_rounds'], d['huge_components'])"; done
for r in 64 256 1024; do PCB_HUGE_READS=$r PCB_BENCH_CONFIG=cfg4 timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-files-e2e --no-parity 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('reads-threshold $r (cfg4)', d['ms_per_step'], d['stage_ms']['replay'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02k_launches_dense03.csv python tests/scale_check.py dense 0.3 > gpurun_out/ncu_k.log 2>&1; echo "ncu rc=$?"
