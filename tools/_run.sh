mkdir -p gpurun_out/r3i
for f in "" "--full-mesh"; do
python bench.py --skip-methods --skip-e2e --skip-cpu $f > gpurun_out/r3i/bench$f.json 2> gpurun_out/r3i/bench$f.err
tail -c 300 gpurun_out/r3i/bench$f.err; python - <<P
import json
d=json.load(open('gpurun_out/r3i/bench$f.json'))
print("$f", d['ms_per_step'], {k:(round(v['ms'],2), round(v.get('frac_of_hbm',0),3)) for k,v in d['stages'].items()}, d['parity'].get('cross_n'))
P
done
