# the two commands the driver runs, as it runs them
mkdir -p gpurun_out
( time python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02o_reference.json 2> gpurun_out/r02o_reference.err ) 2>&1 | grep real; echo "reference rc=$?"
( time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02o_reference.json; echo; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02o_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e'], d.get('files_e2e'), d.get('cpu_baseline'), d['roofline'], d['clocks'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
