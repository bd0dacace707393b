mkdir -p gpurun_out/r3l
O=gpurun_out/r3l
python bench.py --skip-e2e --skip-methods --skip-cpu > $O/bench_new.json 2> $O/bench_new.err
python - <<P
import json
d=json.load(open('$O/bench_new.json'))
print("new", d['ms_per_step'], {k:(round(v['ms'],2), round(v.get('frac_of_hbm',0),3)) for k,v in d['stages'].items()}, d['parity'].get('cross_n'))
P
