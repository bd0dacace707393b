mkdir -p gpurun_out/r3g
python -m pytest tests -x -q -m gpu -k "fft" > gpurun_out/r3g/pytest_fft.log 2>&1; tail -1 gpurun_out/r3g/pytest_fft.log
for i in 1 2; do
KB_S=1480 python tools/kbench.py fft 2>&1 | grep "series-major T=131072" | sed 's/^/new /' | tee -a gpurun_out/r3g/kb.log
KB_S=1480 DYNAPHOPY_B200_LIB=tools/_variants/libdp_dev.so python tools/kbench.py fft 2>&1 | grep "series-major T=131072" | sed 's/^/old /' | tee -a gpurun_out/r3g/kb.log
done
