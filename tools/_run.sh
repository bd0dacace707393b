mkdir -p gpurun_out/r4b
for i in 1 2; do
KB_S=1480 python tools/kbench.py fft 2>&1 | grep "series-major T=131072" | sed 's/^/base /' | cut -c1-120 | tee -a gpurun_out/r4b/kb.log
KB_S=1480 DYNAPHOPY_B200_LIB=tools/_variants/libdp_dev.so python tools/kbench.py fft 2>&1 | grep "series-major T=131072" | sed 's/^/ld128 /' | cut -c1-120 | tee -a gpurun_out/r4b/kb.log
done
DYNAPHOPY_B200_LIB=tools/_variants/libdp_dev.so python -m pytest tests/test_parity.py -x -q -m gpu -k fft 2>&1 | tail -1
