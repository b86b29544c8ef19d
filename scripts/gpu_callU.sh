# giant tier: resident warps per SM (the exec kernel waits on global loads: 34 stall cycles per issue, ncu r02ae)
for w in 24 32 40 48; do for cap in 2048 1024; do
PCB_HUGE_WARPS=$w PCB_HUGE_CAP=$cap timeout 300 python tests/scale_check.py dense 0.3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('huge_warps $w cap $cap dense0.3 replay', d['stage_ms']['replay'], 'rounds', d['huge_rounds'])"
done; done
