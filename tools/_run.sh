mkdir -p gpurun_out/r4c
O=gpurun_out/r4c
python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py > $O/bench_g1.json 2> $O/bench_g1.err
tail -c 300 $O/bench_g1.err
python - <<P
import json
d=json.load(open('$O/bench_g1.json'))
print(d['ms_per_step'], {k:(round(v['ms'],2), round(v.get('frac_of_hbm',0),3)) for k,v in d['stages'].items()}, 'e2e', d['e2e']['ms_per_step'], d['parity']['ok'], d['roofline']['frac'], d['roofline']['fp64']['frac'], d['clocks'])
P
