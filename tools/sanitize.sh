#!/bin/bash
# compute-sanitizer runs of the hot-path kernels on small shapes (tools/sanitize_case.py); logs -> $1 (default
# gpurun_out/sanitize).  Usage on a GPU box:  bash tools/sanitize.sh [outdir] [cases...]
out=${1:-gpurun_out/sanitize}; shift
cases=${@:-fft fft_big project dense mem_cluster corr}
mkdir -p "$out"
for c in $cases; do
  for tool in memcheck racecheck synccheck; do
    log="$out/${c}_${tool}.log"
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py $c > "$log" 2>&1
    echo "$c $tool exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tr '\n' ' ')" | tee -a "$out/summary.txt"
  done
done
