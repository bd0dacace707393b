mkdir -p gpurun_out/r3t
O=gpurun_out/r3t
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py --skip-e2e --skip-methods --skip-cpu > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=json.load(open('$O/bench.json'))
print(d['ms_per_step'], {k:(round(v['ms'],2), round(v.get('frac_of_hbm',0),3)) for k,v in d['stages'].items()}, d['parity'].get('cross_n'))
P
