N=$1
mkdir -p gpurun_out/r3r
for v in 111 128 series; do
export DP_CYCLIC_TRACE=1
unset DP_NO_VC_EXCHANGE DP_SPEC_CTAS
[ $v = series ] && export DP_NO_VC_EXCHANGE=1 || export DP_SPEC_CTAS=$v
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --skip-e2e --skip-methods > gpurun_out/r3r/bench_g${N}_$v.json 2> gpurun_out/r3r/bench_g${N}_$v.err
tail -c 400 gpurun_out/r3r/bench_g${N}_$v.err | grep -v "OMP\|\*\*\*"
grep "cyclic trace" gpurun_out/r3r/bench_g${N}_$v.json | tail -1; python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r3r/bench_g${N}_$v.json') if l.startswith('{')][-1])
print("variant $v:", d['ms_per_step'], {k:(round(v['ms'],2) if v.get('ms') is not None else None) for k,v in d['stages'].items() if isinstance(v,dict) and 'ms' in v}, d['parity'].get('cross_n',{}).get('max_rel_err_series_sums'))
P
done
