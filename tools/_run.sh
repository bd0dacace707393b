mkdir -p gpurun_out/r3y
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu -s > gpurun_out/r3y/pytest_multigpu.log 2>&1
grep -E "FAIL|PARITY|passed|failed|Error|exchanged" gpurun_out/r3y/pytest_multigpu.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --skip-methods > gpurun_out/r3y/bench_g2.json 2> gpurun_out/r3y/bench_g2.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r3y/bench_g2.json') if l.startswith('{')][-1])
print(d['ms_per_step'], {k:(round(v['ms'],2) if v.get('ms') is not None else None) for k,v in d['stages'].items() if isinstance(v,dict) and 'ms' in v}, d['parity'].get('cross_n',{}).get('max_rel_err_series_sums'), 'e2e', d.get('e2e',{}).get('ms_per_step'))
P
CUDA_VISIBLE_DEVICES=0 python -m pytest tests -x -q -m gpu > gpurun_out/r3y/pytest_gpu.log 2>&1; tail -1 gpurun_out/r3y/pytest_gpu.log
