"""Condense an `ncu --set full` report into the few numbers DESIGN.md / profiles/ quote.

usage: python tools/ncu_summary.py report.ncu-rep [out.json]
Prints one JSON object per profiled launch: duration, DRAM bytes, L2 hit rate, pipe utilisation,
registers, shared memory, top stall reasons and the hottest SASS lines (needs -lineinfo builds)."""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_inst_pct",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_cycles_active_pct",
    "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg": "dmma_cycles_active_avg",
    "smsp__cycles_active.avg": "smsp_cycles_active_avg",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_inst_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__block_size": "block",
    "launch__grid_size": "grid",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "smsp__inst_executed.sum": "warp_instructions",
    "sass__inst_executed_local_loads": "local_loads",
    "sass__inst_executed_local_stores": "local_stores",
}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                d[name] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    return res


def source_page(rep, topn=8):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return {}
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0 for h in stall_cols}
    lines, nsamp = [], 0
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        try:
            n = int(r[idx["# Samples"]])
        except ValueError:
            continue
        nsamp += n
        for h in stall_cols:
            tot[h] += int(r[idx[h]] or 0)
        lines.append((n, r[idx["Source"]].strip()))
    lines.sort(key=lambda x: -x[0])
    return {"samples": nsamp,
            "stalls_pct": {h[6:]: round(100.0 * v / max(nsamp, 1), 1) for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:6]},
            "hottest_sass": [f"{100.0 * n / max(nsamp, 1):.1f}% {s[:60]}" for n, s in lines[:topn]]}


if __name__ == "__main__":
    rep = sys.argv[1]
    res = raw_page(rep)
    if res:
        res[0].update(source_page(rep))
    text = "\n".join(json.dumps(d) for d in res)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
