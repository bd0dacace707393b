mkdir -p gpurun_out/r3f
python -m pytest tests -x -q -m gpu > gpurun_out/r3f/pytest_gpu.log 2>&1
tail -3 gpurun_out/r3f/pytest_gpu.log
python bench.py --workload cfg5-B > gpurun_out/r3f/bench_5b.json 2> gpurun_out/r3f/bench_5b.err
tail -c 600 gpurun_out/r3f/bench_5b.err; head -c 3000 gpurun_out/r3f/bench_5b.json
