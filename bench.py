#!/usr/bin/env python
"""Benchmark of the normal-mode-decomposition hot path on BASELINE.json's headline workload.

Workload (BASELINE.json configs[4], SURVEY.md section 8d "cfg 5-A"): synthetic 2-atom cubic cell,
supercell 16x16x20 = 10 240 atoms, 2^17 = 131 072 frames (dt = 1 fs), the full commensurate q-mesh
(5120 q, 6 modes each = 30 720 series), default frequency grid 0..40 THz step 0.05 (FFT windows of
20 000 samples, 7 windows).  One step = wave-vector projection of every frame onto every q
(A1+A2) -> eigenvector contraction (A3) -> [all-to-all when N > 1] -> FFT power spectra of every
series (A7).  The total problem is fixed and the frames are sharded over the ranks by time block
("scaling": "strong").  N > 1: the contracted chunks travel over NVLink peer memory (symmetric memory, copy
engines; NCCL all-to-all when that is unavailable); for N >= 4 the 4096-frame blocks are dealt to the ranks
cyclically so that the spectra of finished windows overlap the rest of the exchange (--sharding).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, sm_100a)
    python bench.py --impl reference [...]                         # reference arm (host CPU)

metric/value: atom*frames per second through the whole step, all ranks together (inputs resident
in HBM).  e2e: the same through MeshPipeline.run_from_host with the trajectory in pinned HOST
memory (H2D of every frame and D2H of the spectra inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ---- workload definition ------------------------------------------------------------
DIMS = (16, 16, 20)
A_LAT = 4.0
UNIT_FRAC = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
UNIT_MASS = np.array([28.0855, 69.723])
DT = 0.001
LINES = ((3.83, 1.0), (7.11, 0.7), (15.53, 0.4))
KB = 0.831446            # u A^2 / ps^2 / K  (reference's k_B, iofile/__init__.py:348)
TEMP = 400.0
SEED = 20260101
FREQ = np.arange(0, 40.05, 0.05)
WORKLOAD = "cfg5-A"      # --workload cfg5-B: ONE window of the whole run (resolution 1/(T dt)), 5244 frequencies up to 40 THz
CHUNK = 1024             # frames generated / copied per chunk


FP64_PEAK_TOPS = 18.3     # measured fp64 lane-ops/s (tools/micro/fp64_rate.cu); x2 flop for FMA


def workload(frames):
    cell = np.eye(3) * A_LAT
    nc = int(np.prod(DIMS))
    unit_pos = UNIT_FRAC @ cell
    grid = np.array([[x, y, z] for z in range(DIMS[2]) for y in range(DIMS[1]) for x in range(DIMS[0])], float) @ cell
    pos = (unit_pos[:, None, :] + grid[None]).reshape(-1, 3)
    typ = np.repeat(np.arange(2), nc).astype(np.int32)
    mass = np.repeat(UNIT_MASS, nc)
    m = np.array([[mx, my, mz] for mz in range(DIMS[2]) for my in range(DIMS[1]) for mx in range(DIMS[0])])
    q_red = m / np.array(DIMS)                                 # identity primitive matrix
    q_cart = q_red @ (2.0 * np.pi * np.linalg.inv(cell).T)     # Quasiparticle.get_q_vector
    return dict(cell=cell, nc=nc, pos=pos, typ=typ, mass=mass, q_cart=q_cart, n_atoms=2 * nc, frames=frames,
                n_prim=2, Q=len(q_cart), M=6)


def gen_chunk_torch(torch, dev, wl, t0, t1):
    """Frames [t0, t1) of the synthetic trajectory (white noise + three spectral lines); the random
    stream depends only on the chunk index, so the data are the same for every rank count."""
    assert t0 % CHUNK == 0
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + t0 // CHUNK)
    n = wl["n_atoms"]
    sigma = torch.as_tensor(np.sqrt(KB * TEMP / wl["mass"]), device=dev)[None, :, None]
    v = torch.randn((t1 - t0, n, 3), dtype=torch.float64, device=dev, generator=g)
    gp = torch.Generator(device=dev)
    gp.manual_seed(SEED - 1)
    t = (torch.arange(t0, t1, device=dev, dtype=torch.float64) * DT)[:, None, None]
    for f0, a0 in LINES:
        phi = torch.rand((1, n, 3), dtype=torch.float64, device=dev, generator=gp) * (2 * np.pi)
        v += a0 * torch.cos(2 * np.pi * f0 * t + phi)
    return v * sigma


def random_unitary_eigenvectors(torch, dev, Q, M):
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + 7)
    a = torch.randn((Q, M, M), dtype=torch.float64, device=dev, generator=g) + \
        1j * torch.randn((Q, M, M), dtype=torch.float64, device=dev, generator=g)
    qmat, _ = torch.linalg.qr(a)
    return qmat.transpose(1, 2).contiguous()                   # e[q][s][j] = E[j][s]


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except Exception:
                continue
            for nme, val in zip(names, parts[2:]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ---- CPU baseline: the reference's own algorithm (oracle port / compiled reference) on host cores ----
def cpu_sample(wl, v_sub, eig_sub, q_idx, n_series, frames_total):
    """Bounded sample of the same workload on the host: projection of v_sub (T_sub frames, all atoms)
    onto len(q_idx) q, contraction, and FFT spectra of n_series full-length series; linear
    extrapolation to the whole job (projection ~ T*Q, spectra ~ number of series)."""
    from oracle import port
    vm = port.velocity_mass_average(v_sub.astype(complex), wl["mass"])     # reference arrays are complex128
    t0 = time.perf_counter()
    vcs = [port.project_onto_wave_vector(vm, wl["pos"], wl["typ"], wl["n_prim"], wl["q_cart"][q]) for q in q_idx]
    t_proj = time.perf_counter() - t0
    t0 = time.perf_counter()
    vqs = [port.project_onto_phonon(vc, eig_sub[i].reshape(6, 2, 3)) for i, vc in enumerate(vcs)]
    t_ph = time.perf_counter() - t0
    rng = np.random.default_rng(SEED)
    series = rng.normal(size=(frames_total, n_series)) + 1j * rng.normal(size=(frames_total, n_series))
    t0 = time.perf_counter()
    port.fft_spectra(series, DT, FREQ)                                      # np.correlate + np.fft, as the reference
    t_fft = time.perf_counter() - t0
    t_sub = v_sub.shape[0]
    scale_proj = (frames_total / t_sub) * (wl["Q"] / len(q_idx))
    t_full = (t_proj + t_ph) * scale_proj + t_fft * (wl["Q"] * wl["M"] / n_series)
    return {"t_proj_sample_s": t_proj, "t_phonon_sample_s": t_ph, "t_fft_sample_s": t_fft,
            "t_full_extrapolated_s": t_full, "value": wl["n_atoms"] * frames_total / t_full, "_vq": vqs}


_REF = {}          # data of the pooled reference sample (inherited by the forked workers, copy on write)


def _ref_task(task):
    """One unit of the pooled reference sample: the projection + contraction of the sampled frames for ONE wave vector
    (what one pass of the q loop of get_commensurate_points_data does), or the FFT spectrum of ONE series."""
    from oracle import port
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=1)
    except Exception:                                   # no threadpoolctl: BLAS keeps its default
        limit = None
    kind, i = task
    wl = _REF["wl"]
    if kind == "proj":
        vc = port.project_onto_wave_vector(_REF["vm"], wl["pos"], wl["typ"], wl["n_prim"], wl["q_cart"][_REF["q_idx"][i]])
        port.project_onto_phonon(vc, _REF["eig"][i].reshape(6, 2, 3))
    else:
        port.fft_spectra(_REF["series"][:, i:i + 1], DT, FREQ)
    if limit is not None:
        limit.restore_original_limits() if hasattr(limit, "restore_original_limits") else None
    return 0


def cpu_sample_pooled(wl, v_sub, n_q, n_series, frames_total, procs):
    """The reference's arithmetic (oracle port) on ALL host cores: the q loop and the loop over series are
    embarrassingly parallel, so a process pool takes one wave vector / one series per task.  Wall times of the two
    pools, extrapolated linearly to the whole job like cpu_sample."""
    import multiprocessing as mp
    from oracle import port
    if "vm" not in _REF:                                  # inputs of the sample: built once, outside the timed pools
        rng = np.random.default_rng(SEED)
        _REF["wl"] = wl
        _REF["vm"] = port.velocity_mass_average(v_sub.astype(complex), wl["mass"])
        _REF["q_idx"] = [int(q) for q in np.linspace(0, wl["Q"] - 1, n_q).astype(int)]
        _REF["eig"] = np.stack([np.linalg.qr(rng.normal(size=(6, 6)) + 1j * rng.normal(size=(6, 6)))[0].T for _ in range(n_q)])
        _REF["series"] = rng.normal(size=(frames_total, n_series)) + 1j * rng.normal(size=(frames_total, n_series))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_ref_task, [("fft", 0)] * procs)                      # workers started and warm before the clock
        t0 = time.perf_counter()
        pool.map(_ref_task, [("proj", i) for i in range(n_q)], chunksize=1)
        t_proj = time.perf_counter() - t0
        t0 = time.perf_counter()
        pool.map(_ref_task, [("fft", i) for i in range(n_series)], chunksize=1)
        t_fft = time.perf_counter() - t0
    t_sub = v_sub.shape[0]
    t_full = t_proj * (frames_total / t_sub) * (wl["Q"] / n_q) + t_fft * (wl["Q"] * wl["M"] / n_series)
    return {"t_proj_sample_s": t_proj, "t_fft_sample_s": t_fft, "t_full_extrapolated_s": t_full,
            "value": wl["n_atoms"] * frames_total / t_full}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = args.ref_procs if args.ref_procs > 0 else (os.cpu_count() or 1)
    if procs > 1:
        return run_reference_pooled(args, procs)
    frames = args.frames
    wl = workload(frames)
    rng = np.random.default_rng(SEED)
    t_sub = args.cpu_frames
    sigma = np.sqrt(KB * TEMP / wl["mass"])[None, :, None]
    v_sub = rng.normal(size=(t_sub, wl["n_atoms"], 3)) * sigma             # white noise of the workload's amplitude
    q_idx = [0, 1237, 2560, 5119][: args.cpu_q]
    eig = np.stack([np.linalg.qr(rng.normal(size=(6, 6)) + 1j * rng.normal(size=(6, 6)))[0].T for _ in q_idx])
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_sample(wl, v_sub, eig, q_idx, args.cpu_series, frames)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    value = res["value"]
    sample = (f"projection of {t_sub} frames x {wl['n_atoms']} atoms onto {len(q_idx)} q + contraction, "
              f"FFT spectra (np.correlate + np.fft) of {args.cpu_series} series of {frames} samples; "
              "scaled linearly to 131072 frames x 5120 q and 30720 series")
    line = {"impl": "reference", "metric": "atom_frames_per_s", "value": value, "unit": "atom*frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl, frames, args.gpus),
            "cpu_baseline": {"value": value, "unit": "atom*frames/s", "cores": 1, "kind": "port",
                             "sample": sample, "host_cores": os.cpu_count(),
                             "note": "the reference's FFT-method path is serial numpy (per-atom np.exp products, np.correlate, "
                                     "pocketfft; no OpenMP on this path, SURVEY 0.1): one host thread does the work even with "
                                     "BLAS threading left at its default; each step = the bounded sample, value extrapolated",
                             "full_job_extrapolated_s": res["t_full_extrapolated_s"]},
            "e2e": {"value": value, "unit": "atom*frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_pooled(args, procs):
    frames = args.frames
    wl = workload(frames)
    rng = np.random.default_rng(SEED)
    t_sub = args.cpu_frames
    sigma = np.sqrt(KB * TEMP / wl["mass"])[None, :, None]
    v_sub = rng.normal(size=(t_sub, wl["n_atoms"], 3)) * sigma             # white noise of the workload's amplitude
    n_q, n_series = max(args.cpu_q, procs), max(args.cpu_series, 2 * procs)
    times, res = [], None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_sample_pooled(wl, v_sub, n_q, n_series, frames, procs)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    value = res["value"]
    sample = (f"{procs} worker processes (one per host core, BLAS limited to one thread each): projection of {t_sub} frames "
              f"x {wl['n_atoms']} atoms onto {n_q} q + contraction ({res['t_proj_sample_s']:.2f} s wall), FFT spectra "
              f"(np.correlate + np.fft) of {n_series} series of {frames} samples ({res['t_fft_sample_s']:.2f} s wall); "
              "scaled linearly to 131072 frames x 5120 q and 30720 series")
    line = {"impl": "reference", "metric": "atom_frames_per_s", "value": value, "unit": "atom*frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl, frames, args.gpus),
            "cpu_baseline": {"value": value, "unit": "atom*frames/s", "cores": procs, "kind": "port",
                             "sample": sample, "host_cores": os.cpu_count(),
                             "note": "the reference's FFT-method path is serial numpy (per-atom np.exp products, np.correlate, "
                                     "pocketfft; no OpenMP on this path, SURVEY 0.1); here its q loop and its loop over series "
                                     "are spread over all host cores with a process pool -- the most a user can get out of the "
                                     "reference on this box without changing its arithmetic (--ref-procs 1: one thread, as "
                                     "shipped); each step = the bounded sample, value extrapolated",
                             "full_job_extrapolated_s": res["t_full_extrapolated_s"]},
            "e2e": {"value": value, "unit": "atom*frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_dict(wl, frames, gpus):
    spectra = ("801 frequencies, 7 windows of 20000" if WORKLOAD == "cfg5-A" else
               f"{len(FREQ)} frequencies at the resolution 1/(T dt): one window of {frames} samples, padded transforms of "
               f"{2 * frames} points")
    return {"workload": f"{WORKLOAD} synthetic 2-atom cubic cell, 16x16x20 supercell (10240 atoms), "
                        f"{frames} frames, full commensurate q-mesh (5120 q x 6 modes), "
                        f"projection + eigenvector contraction + FFT power spectra ({spectra})",
            "atoms": wl["n_atoms"], "frames": frames, "q_points": wl["Q"], "modes_per_q": wl["M"],
            "series": wl["Q"] * wl["M"], "frequencies": len(FREQ), "dt_ps": DT,
            "intermediate": "pair-major bare lattice DFT; one wave vector per pair (q, -q) is projected and stored (2564 of 5120), both are contracted (--full-mesh: all 5120)",
            "sharding": f"time-block x{gpus}" + (" (4096-frame blocks dealt cyclically to the ranks) + one exchange over NVLink "
                                                   "peer memory (copy engines; NCCL all-to-all fallback); N >= 4: the ranks exchange "
                                                   "the half-mesh intermediate and keep their spectra in their own order of wave "
                                                   "vectors (natural order restored after the timed region for the parity check)"
                                                   if gpus > 1 else ""),
            "l2_policy": "inputs (32 GB velocities, 64 GB projected series) far exceed the 126 MB L2"}


def ncu_traffic_from_profiles(units_now):
    """stage -> DRAM bytes of this launch, from profiles/ncu_index.json: a list of
    {"stage", "file", "kernel", "units"} entries (newest last) pointing at tools/ncu_summary.py outputs."""
    out, src = {}, {}
    idx_path = os.path.join(ROOT, "profiles", "ncu_index.json")
    if not os.path.exists(idx_path):
        return out, src
    for ent in json.load(open(idx_path)):
        path = os.path.join(ROOT, "profiles", ent["file"])
        if ent["stage"] not in units_now or not os.path.exists(path):
            continue
        for line in open(path):
            try:
                d = json.loads(line)
            except ValueError:
                continue
            if ent["kernel"] in d.get("kernel", "") and "dram_read" in d:
                def gb(x):
                    val, unit = x.split()[:2]
                    return float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit]
                out[ent["stage"]] = (gb(d["dram_read"]) + gb(d["dram_write"])) / ent["units"] * units_now[ent["stage"]]
                src[ent["stage"]] = "profiles/" + ent["file"]
                break
    return out, src


def methods_cpu_baselines(methods, send, ps_m_head, frames, torch, ops):
    """CPU side of `spectra_methods` (N = 1, rank 0): the reference's own compiled C kernels (oracle/_ref: c/mem.c and
    c/correlation.c built with -O2 -fopenmp by oracle/Makefile) on the host cores, on a bounded sample of THIS run's
    contracted series; MEM doubles as the parity check of the full-length device result."""
    from oracle import ref
    if not ref.available():
        return
    cores = os.cpu_count()
    # maximum entropy: 8 full-length series, deterministic build (Data[N] := 0)
    x8 = np.ascontiguousarray(send[0, :8, :].t().cpu().numpy())
    t0 = time.perf_counter()
    ref_m = np.stack([ref.mem_det().mem(FREQ, np.ascontiguousarray(x8[:, i]), DT, coefficients=1000) for i in range(8)], axis=1)
    t_m = time.perf_counter() - t0
    ref_m = np.nan_to_num(ref_m) * 0.00010585723
    methods["mem"]["cpu_baseline"] = {"value": 8 / t_m, "unit": "series/s", "kind": "reference", "cores": cores,
                                      "sample": "8 series of %d samples, mem.mem of c/mem.c (OpenMP over frequencies), %.2f s" % (frames, t_m)}
    methods["mem"]["parity"] = {"max_rel_err": float(np.abs(ps_m_head - ref_m[:, :2]).max() / np.abs(ref_m[:, :2]).max()),
                                "peak_bin_matches": int((ps_m_head.argmax(0) == ref_m[:, :2].argmax(0)).sum()), "series": 2,
                                "vs": "compiled reference c/mem.c (deterministic build) on the same series"}
    # direct correlation: as shipped, N = 4096 (the reference recomputes the lag products for every frequency: O(F N^2 / step))
    n_c = 4096
    xc = np.ascontiguousarray(x8[:n_c, :2])
    t0 = time.perf_counter()
    for i in range(2):
        ref.correlation_shipped().correlation_par(FREQ, np.ascontiguousarray(xc[:, i]), DT, step=10, integration_method=1)
    t_c = time.perf_counter() - t0
    # parity of the device kernel against the reference's intended-weight build on the same truncated series
    ref_c = ref.correlation_intended().correlation_par(FREQ, np.ascontiguousarray(xc[:, 0]), DT, step=10, integration_method=1) * 0.00010585723
    dev_c = ops.power_correlation(send[0, :1, :n_c].contiguous(), DT, FREQ, 10, 1, 1, series_major=True)[:, 0].cpu().numpy()
    methods["correlation"]["parity"] = {"max_rel_err": float(np.abs(dev_c - ref_c).max() / np.abs(ref_c).max()),
                                        "peak_bin_matches": int(dev_c.argmax() == ref_c.argmax()), "series": 1,
                                        "vs": "compiled reference c/correlation.c (intended-weight build), first %d samples" % n_c}
    per_series_full = t_c / 2 * (frames / n_c) ** 2
    methods["correlation"]["cpu_baseline"] = {"value": 1.0 / per_series_full, "unit": "series/s", "kind": "reference", "cores": cores,
                                              "sample": "2 series of %d samples, correlation.correlation_par of c/correlation.c (OpenMP "
                                                        "over frequencies), %.2f s; scaled by N^2 to %d samples (extrapolated)" % (n_c, t_c, frames)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from dynaphopy_b200 import ops as dops, pipeline
    ops = dops.Ops()
    frames = args.frames
    assert frames % (world * CHUNK) == 0, "frames must be a multiple of ranks * 1024"
    wl = workload(frames)
    tl = frames // world
    t_lo = rank * tl
    plan = pipeline.CommensuratePlan(wl["cell"], DIMS, wl["pos"], wl["mass"], wl["typ"], wl["n_prim"], wl["q_cart"])
    assert plan.usable() and plan.commensurate.all()
    eig = random_unitary_eigenvectors(torch, dev, wl["Q"], wl["M"])
    pipe = pipeline.MeshPipeline(plan, eig, FREQ, DT, ops=ops, algorithm=2, use_mirror=not args.full_mesh)
    assert pipe.fold_twiddle        # primitive cell: the projection twiddle rides on the eigenvectors

    # ---- device-resident inputs -------------------------------------------------------------------
    CH = args.chunk                                  # frames per projection/contraction chunk (4096: vc chunk of 2 GB)
    # N > 1: chunk-cyclic time sharding (chunk j of the trajectory lives on rank j % N), so that the spectra of finished
    # windows overlap the rest of the exchange; --sharding blocks keeps one contiguous time block per rank
    # auto: cyclic where the exchange is the longer leg (measured: 8 GPUs 27.9 vs 32.3 ms, 4 GPUs equal, 2 GPUs 103.9 vs 99.6)
    cyclic = world > 1 and (args.sharding == "cyclic" or (args.sharding == "auto" and world >= 4))
    v = torch.empty((tl, wl["n_atoms"], 3), dtype=torch.float64, device=dev)
    if cyclic:
        assert tl % CH == 0
        for c, (g0, g1) in enumerate(pipe.cyclic_chunks(frames, CH)):
            for t0 in range(0, CH, CHUNK):
                v[c * CH + t0:c * CH + t0 + CHUNK] = gen_chunk_torch(torch, dev, wl, g0 + t0, g0 + t0 + CHUNK)
    else:
        for t0 in range(0, tl, CHUNK):
            v[t0:t0 + CHUNK] = gen_chunk_torch(torch, dev, wl, t_lo + t0, t_lo + t0 + CHUNK)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage_ms = {"project": [], "contract": [], "exchange": [], "spectra": []}
    span_ms = []
    use_peers = peer_stores = False
    if world == 1:
        send = torch.empty((1, pipe.q_block * pipe.M, tl), dtype=torch.complex128, device=dev)
    elif cyclic:
        send = send2 = recv = None
        use_peers = pipe.peer_exchange_setup(tl // CH, CH)
    else:
        assert tl % CH == 0
        send = None
        # preferred: the contraction kernel stores every block straight into the destination rank's receive buffer
        # (symmetric memory over NVLink); fallback: two send slots + chunked asynchronous NCCL all-to-all
        use_peers = pipe.peer_exchange_setup(tl // CH, CH)
        peer_stores = use_peers and pipe.peer_mode() == "stores"
        if use_peers:
            recv = pipe._peer["recv"]
            peer_ptrs = pipe._peer["ptrs"]
            send2 = None if peer_stores else torch.empty((2, world, pipe.q_block * pipe.M, CH), dtype=torch.complex128, device=dev)
        else:
            send2 = torch.empty((2, world, pipe.q_block * pipe.M, CH), dtype=torch.complex128, device=dev)
            recv = torch.empty((world, tl // CH, pipe.q_block * pipe.M, CH), dtype=torch.complex128, device=dev)

    def step(record):
        """One pass of the hot path over the rank's frames: [project chunk -> contract chunk into the
        series-major send buffer (-> asynchronous all-to-all of the chunk when N > 1)]* -> FFT spectra.
        Stage times are CUDA-event sums; with N > 1 the exchange runs on NCCL's stream underneath the
        projection of the following chunks and "exchange" is only what remains exposed at the end."""
        if cyclic:
            rec = {}
            # the spectra stay with the rank that contracted them, in its order of wave vectors (local_q); they are put
            # into the natural order after the timed region, where the parity block gathers them
            ps = pipe.run_cyclic(v, CH, record=rec if record else None, natural_order=False)
            if record:
                for k in stage_ms:
                    stage_ms[k].append(rec.get(k, 0.0))
                span_ms.append(rec.get("span", 0.0))
            return ps
        evs = []
        pending = [None, None]
        if world > 1 and use_peers:
            pipe.peer_begin()                        # device barrier: the peers have finished reading the previous step's blocks
        for ci, t0 in enumerate(range(0, tl, CH)):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            vc = pipe.project_chunk(v[t0:t0 + CH])
            b.record()
            if world == 1:
                pipe.contract_chunk(vc, out=send, t_offset=t0, t_total=tl)
                c.record()
            elif use_peers and peer_stores:              # the contraction kernel stores into the peers' receive buffers
                off = ((rank * (tl // CH) + ci) * pipe.q_block * pipe.M * CH) * 16
                pipe.contract_chunk(vc, t_total=CH, block_ptrs=[peer_ptrs[dst] + off for dst in range(world)])
                c.record()
            elif use_peers:                              # local send slot, copy engines push the blocks over NVLink
                slot = ci % 2
                pipe.peer_wait_slot(slot)
                pipe.contract_chunk(vc, t_total=CH, block_ptrs=pipe.peer_block_ptrs(send2[slot], ci))
                c.record()
                pipe.peer_push(send2[slot], ci, slot)
            else:
                slot = ci % 2
                if pending[slot] is not None:
                    pending[slot].wait()
                pipe.contract_chunk(vc, out=send2[slot], t_total=CH)
                c.record()
                outs = [torch.view_as_real(recv[src, ci]) for src in range(world)]
                ins = [torch.view_as_real(send2[slot, dst]) for dst in range(world)]
                pending[slot] = dist.all_to_all(outs, ins, async_op=True)
            del vc
            evs.append((a, b, c))
        e2 = torch.cuda.Event(enable_timing=True); e2.record()
        for w in pending:
            if w is not None:
                w.wait()
        if world > 1 and use_peers:
            pipe.peer_end()                          # own copies done + device barrier: every rank's blocks have landed
        blocks = send if world == 1 else recv.reshape(world * (tl // CH), pipe.q_block * pipe.M, CH)
        e3 = torch.cuda.Event(enable_timing=True); e3.record()
        ps = pipe.spectra_series(blocks)
        e4 = torch.cuda.Event(enable_timing=True); e4.record()
        if record:
            torch.cuda.synchronize()
            stage_ms["project"].append(sum(a.elapsed_time(b) for a, b, c in evs))
            stage_ms["contract"].append(sum(b.elapsed_time(c) for a, b, c in evs))
            stage_ms["exchange"].append(e2.elapsed_time(e3))
            stage_ms["spectra"].append(e3.elapsed_time(e4))
        return ps

    for _ in range(args.warmup):
        ps = step(False)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    l0 = ops.launch_count()
    w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ps = step(True)
    e1.record()
    barrier()
    w1 = time.time()
    launches = ops.launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = wl["n_atoms"] * frames / (ms_per_step * 1e-3)
    if cyclic and getattr(pipe, "_vcx", None):
        ps = pipe.restore_natural_order(ps)           # (outside the timed region: see step())
    peak_bin = int(torch.argmax(ps[:, 0]).item())
    checksum = float(ps.sum().item())
    # per-series sums of the spectra, in global series order (rank g owns the series of q block g): the cross-N parity
    # reference -- every rank count must reproduce the single-GPU values (tests/golden/bench_cfg5a_series_sums.npy)
    series_sums = ps.sum(dim=0)
    series_peaks = torch.argmax(ps, dim=0).to(torch.float64)
    if world > 1:
        gathered = [torch.empty_like(series_sums) for _ in range(world)]
        dist.all_gather(gathered, series_sums)
        series_sums = torch.cat(gathered)
        gathered = [torch.empty_like(series_peaks) for _ in range(world)]
        dist.all_gather(gathered, series_peaks)
        series_peaks = torch.cat(gathered)
    series_sums = series_sums.cpu().numpy()
    series_peaks = series_peaks.cpu().numpy().astype(np.int64)
    parity = {"tolerance": 1e-10}
    gpath = os.path.join(ROOT, "tests", "golden", "bench_cfg5a_series_sums.npz")
    if args.write_golden and world == 1 and rank == 0 and frames == 131072 and WORKLOAD == "cfg5-A":
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "bench_cfg5a_series_sums.npz"), sums=series_sums, peaks=series_peaks)
    if os.path.exists(gpath) and frames == 131072 and WORKLOAD == "cfg5-A":
        gold = np.load(gpath)
        parity["cross_n"] = {"reference": "single-GPU run of this bench (tests/golden/bench_cfg5a_series_sums.npz)",
                             "series": int(series_sums.size),
                             "max_rel_err_series_sums": float(np.abs(series_sums - gold["sums"]).max() / np.abs(gold["sums"]).max()),
                             "peak_bin_matches": int((series_peaks == gold["peaks"]).sum())}
    # oracle parity on a sample of THIS run's data (world == 1: the contracted series are all on this device)
    if world == 1 and not args.skip_cpu:
        from oracle import port as _port
        q_idx_p = [0, 1237, 2560, 5119]
        tp = 1024
        vsub_p = gen_chunk_torch(torch, dev, wl, 0, CHUNK)[:tp].cpu().numpy()
        vm_p = _port.velocity_mass_average(vsub_p.astype(complex), wl["mass"])
        eig_p = eig[q_idx_p].cpu().numpy()
        err_vq = 0.0
        for i, q in enumerate(q_idx_p):
            vc_ref = _port.project_onto_wave_vector(vm_p, wl["pos"], wl["typ"], wl["n_prim"], wl["q_cart"][q])
            vq_ref = _port.project_onto_phonon(vc_ref, eig_p[i].reshape(6, 2, 3))
            vq_dev = send[0, q * wl["M"]:(q + 1) * wl["M"], :tp].t().cpu().numpy()
            err_vq = max(err_vq, float(np.abs(vq_dev - vq_ref).max() / np.abs(vq_ref).max()))
        s_idx = [0, 1, 7427, 15360, 15365, 22222, 30718, 30719]
        ser = send[0, s_idx, :].t().cpu().numpy()                                 # (T, 8) device-contracted series
        ps_ref = _port.fft_spectra(ser, DT, FREQ, use_fft_correlate=True) if frames >= 20000 else _port.fft_spectra(ser, DT, FREQ)
        ps_dev = ps[:, s_idx].cpu().numpy()
        parity["oracle"] = {
            "vq_max_rel_err": err_vq, "vq_sample": "%d q x %d frames x %d modes vs oracle/port.py projection + contraction" % (len(q_idx_p), tp, wl["M"]),
            "fft_spectra_max_rel_err": float(np.abs(ps_dev - ps_ref).max() / np.abs(ps_ref).max()),
            "fft_spectra_sample": "%d series of %d samples vs oracle/port.py (np.correlate restated through FFTs + np.fft + np.interp)" % (len(s_idx), frames),
            "peak_bin_matches": int((ps_dev.argmax(0) == ps_ref.argmax(0)).sum()), "peak_bin_total": len(s_idx)}
        parity["ok"] = bool(err_vq < 1e-10 and parity["oracle"]["fft_spectra_max_rel_err"] < 1e-10
                            and parity["oracle"]["peak_bin_matches"] == len(s_idx))
    if "cross_n" in parity:
        ok_n = parity["cross_n"]["max_rel_err_series_sums"] < 1e-10 and parity["cross_n"]["peak_bin_matches"] == parity["cross_n"]["series"]
        parity["ok"] = bool(parity.get("ok", True) and ok_n)

    # ---- the other two spectrum methods of the selector on a bounded sample of the SAME series (not in the timed
    # region; the metric names "mode spectra/s" per algorithm): maximum entropy (m = 1000) and direct correlation
    # ---- the other two spectrum methods of the selector on the WHOLE job (all 30 720 series, every rank its own share,
    # read in place from the exchanged / contracted blocks -- dp_power_mem_strided, dp_power_correlation_strided): maximum
    # entropy (m = 1000, c/mem.c) and direct correlation (step 10, method 1, intended weights exp(i w tau), c/correlation.c)
    methods = None
    if not args.skip_methods:
        if world == 1:
            blocks = send
        elif cyclic:
            if getattr(pipe, "_vcx", None):
                blocks = pipe._vcx_series                  # (1, S_loc, T): the ranks exchanged the intermediate (local q order)
            else:
                blocks = pipe._peer["cyc_recv"] if use_peers else None
        else:
            blocks = recv.reshape(world * (tl // CH), pipe.q_block * pipe.M, CH)
        if blocks is not None:
            s_all = wl["Q"] * wl["M"]

            def timed(fn):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); out = fn(); b.record(); torch.cuda.synchronize()
                tms = torch.tensor([a.elapsed_time(b) * 1e-3], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
                return float(tms.item()), out
            warm = blocks[:, :64, :]
            ops.power_mem(warm, DT, FREQ, 1000, series_major=True)
            ops.power_correlation(warm, DT, FREQ, 10, 1, 1, series_major=True)
            t_m, ps_m = timed(lambda: ops.power_mem(blocks, DT, FREQ, 1000, series_major=True))
            mem_peak0 = int(torch.argmax(ps_m[:, 0]).item())
            ps_m_head = ps_m[:, :2].cpu().numpy()
            del ps_m
            t_c, ps_c = timed(lambda: ops.power_correlation(blocks, DT, FREQ, 10, 1, 1, series_major=True))
            corr_peak0 = int(torch.argmax(ps_c[:, 0]).item())
            del ps_c
            fl_c = 4.0 * frames * frames / 10 * s_all
            methods = {
                "fft": {"series_per_s": s_all / (np.mean(stage_ms["spectra"]) * 1e-3), "sample": "all series (timed region)"},
                "mem": {"series_per_s": s_all / t_m, "seconds_full_job": t_m, "sample": "all %d series, 1000 coefficients" % s_all,
                        "peak_bin_series0": mem_peak0},
                "correlation": {"series_per_s": s_all / t_c, "seconds_full_job": t_c,
                                "sample": "all %d series, step 10, method 1, intended weights" % s_all,
                                "fp64_TFLOPs": fl_c / t_c / 1e12, "frac_of_fp64_peak": fl_c / t_c / 1e12 / (2 * FP64_PEAK_TOPS * world),
                                "peak_bin_series0": corr_peak0},
            }
            if world == 1 and rank == 0 and not args.skip_cpu:
                methods_cpu_baselines(methods, send, ps_m_head, frames, torch, ops)
            del blocks

    # ---- end to end through the host-facing call (pinned host trajectory) -----------------------------
    e2e = None
    if not args.skip_e2e:
        del v, send
        if world > 1:
            del send2, recv
        del ps
        torch.cuda.empty_cache()
        vh = torch.empty((tl, wl["n_atoms"], 3), dtype=torch.float64, pin_memory=True)
        for t0 in range(0, tl, CHUNK):
            vh[t0:t0 + CHUNK].copy_(gen_chunk_torch(torch, dev, wl, t_lo + t0, t_lo + t0 + CHUNK))
        torch.cuda.synchronize()
        e2e_chunk = 2048 if world == 1 else CH                           # N > 1: the receive-buffer geometry of the timed steps
        pipe.run_from_host(vh, chunk_frames=e2e_chunk)                   # warm-up
        barrier()
        es = max(1, args.steps)
        t0 = time.perf_counter()
        for _ in range(es):
            ps_host = pipe.run_from_host(vh, chunk_frames=e2e_chunk)
        barrier()
        dt_e2e = (time.perf_counter() - t0) / es
        if world > 1:
            tt = torch.tensor([dt_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_e2e = float(tt.item())
        e2e = {"value": wl["n_atoms"] * frames / dt_e2e, "unit": "atom*frames/s", "ms_per_step": dt_e2e * 1e3,
               "h2d_bytes_per_step": int(vh.numel() * 8 * world), "d2h_bytes_per_step": int(ps_host.size * 8 * world),
               "api": "dynaphopy_b200.pipeline.MeshPipeline.run_from_host (pinned host trajectory in, numpy spectra out)",
               "steps": es, "peak_bin_matches_device_run": bool(int(np.argmax(ps_host[:, 0])) == peak_bin)}
        # the host-streamed result against the device-resident run of the timed region: per-series sums, all ranks
        sums_e2e = torch.as_tensor(ps_host.sum(axis=0), device=dev)
        if world > 1:
            gathered = [torch.empty_like(sums_e2e) for _ in range(world)]
            dist.all_gather(gathered, sums_e2e)
            sums_e2e = torch.cat(gathered)
        e2e["max_rel_err_series_sums_vs_device_run"] = float(np.abs(sums_e2e.cpu().numpy() - series_sums).max() / np.abs(series_sums).max())
        del vh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    mean = {k: (statistics.mean(x) if x else 0.0) for k, x in stage_ms.items()}
    # algorithmic bytes per launch on ONE rank
    L, pieces = ops.fft_pieces(frames, DT, FREQ)
    npieces = len({tuple(pc) for pc in pieces})         # a window that spans the whole run is listed twice, read once
    engine = ops.fft_engine(frames, DT, FREQ)
    s_loc = wl["Q"] * wl["M"] // world
    # the intermediate holds one wave vector of every pair (q, -q) (real velocities: Hermitian lattice DFT)
    q_stored = int(pipe.d_qmap.shape[0]) if getattr(pipe, "mirror", False) else wl["Q"]
    bytes_stage = {"project": (24.0 * wl["n_atoms"] + 16.0 * wl["M"] * q_stored) * tl,
                   "contract": 16.0 * wl["M"] * (q_stored + wl["Q"]) * tl,
                   "spectra": (16.0 * L * npieces + 8.0 * len(FREQ)) * s_loc}
    kernels = {"project": "project_fft_kernel", "contract": "project_phonon_series_kernel", "spectra": "spectrum2_kernel" if engine == 2 else "s3_col/s3_row/s3_bins kernels (long-window engine)" if engine == 3 else "spectrum_items_kernel"}
    stages = {}
    for k in ("project", "contract", "spectra"):
        ach = bytes_stage[k] / (mean[k] * 1e-3) / 1e9 if mean[k] > 0 else 0.0
        stages[k] = {"kernel": kernels[k], "ms": mean[k], "achieved_GBps": ach, "frac_of_hbm": ach / hbm}
    stages["exchange"] = {"kernel": ("fused into project_phonon_series_kernel: peer stores over NVLink into the destination's "
                                     "receive buffer (exposed: the closing device barrier)" if (world > 1 and use_peers and peer_stores) else
                                     "copy engines push each contracted chunk into the peers' receive buffers (symmetric memory "
                                     "over NVLink) underneath the following chunks; exposed remainder + device barrier"
                                     if (world > 1 and use_peers) else
                                     "nccl all_to_all per chunk, asynchronous (exposed remainder)") if world > 1 else "none",
                          "ms": mean["exchange"]}
    if cyclic:
        stages["exchange"]["kernel"] = ("copy engines push each contracted chunk into the peers' receive buffers (symmetric memory over "
                                        "NVLink); chunk-cyclic sharding: the spectra of finished windows run on a second stream "
                                        "underneath the later rounds; ms = last contraction end -> last round landed on every rank"
                                        if use_peers else "all_to_all per round (no symmetric memory)")
        if getattr(pipe, "_vcx", None):
            stages["exchange"]["kernel"] = ("the ranks exchange the half-mesh INTERMEDIATE (one wave vector per pair (q, -q): half the bytes "
                                            "of the contracted series): project_fft_kernel writes one block per destination "
                                            "(dp_project_commensurate_blocks), copy engines push them over NVLink peer memory, the "
                                            "destination contracts both members of every pair; chunk-cyclic sharding, spectra of finished "
                                            "windows underneath the later rounds; ms = last projection end -> last round landed on every rank")
        stages["overlap"] = {"span_ms": statistics.mean(span_ms) if span_ms else None,
                             "note": "stage times are CUDA-event sums per stream; spectra overlap projection/contraction/exchange"}
    # DRAM traffic of the dominant kernels: dram__bytes_read.sum + dram__bytes_write.sum of the newest `ncu --set full`
    # capture of each kernel listed in profiles/ncu_index.json (smaller launches of the same kernels: bytes per series /
    # per frame, scaled linearly to this launch)
    ncu_units = {"spectra": s_loc, "project": tl, "contract": tl}
    if engine != 2:
        del ncu_units["spectra"]       # the committed capture is of spectrum2_kernel (cfg 5-A)
    ncu_traffic, ncu_src = ncu_traffic_from_profiles(ncu_units)
    # fp64 view of the spectrum stage: 770.5 k warp-level fp64 instructions per series (7 windows, processed as
    # 3 pairs + 1 single; ncu instruction mix, profiles/r01_ncu_spectrum2.jsonl) against the measured
    # 64 lanes/clk/SM (tools/micro/fp64_rate.cu: 18.3 T lane-ops/s)
    fp64_lane_ops = 770.5e3 * 32 * s_loc if WORKLOAD == "cfg5-A" else 0.0
    fp64 = {"lane_ops": fp64_lane_ops, "achieved_Tops": fp64_lane_ops / (mean["spectra"] * 1e-3) / 1e12 if mean["spectra"] else 0.0,
            "peak_Tops": 18.3, "peak_source": "measured, tools/micro/fp64_rate.cu (DFMA/DADD/DMUL, 63.6 of 64 lanes/clk/SM)"}
    fp64["frac"] = fp64["achieved_Tops"] / fp64["peak_Tops"]
    dom = max(("project", "contract", "spectra"), key=lambda k: mean[k])
    roofline = {"kernel": kernels[dom], "bound": "hbm", "achieved": stages[dom]["achieved_GBps"], "peak": hbm,
                "unit": "GB/s", "frac": stages[dom]["frac_of_hbm"], "traffic": ncu_traffic.get(dom),
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of an ncu --set full capture of a smaller launch "
                                "(%s), scaled linearly to this launch" % ncu_src.get(dom), "algorithmic_bytes": bytes_stage[dom],
                "fp64": fp64 if (dom == "spectra" and WORKLOAD == "cfg5-A") else None, "peak_source": peak_src,
                "share_of_step": mean[dom] / max(sum(mean.values()), 1e-9),
                "note": "algorithmic bytes per launch / CUDA-event duration of the stage on rank 0; "
                        "the FFT-spectrum kernel is fp64-FMA limited before it is HBM limited (DESIGN.md)"}
    line = {"metric": "atom_frames_per_s", "value": value, "unit": "atom*frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(wl, frames, world),
            "mode_spectra_per_s": wl["Q"] * wl["M"] / (ms_per_step * 1e-3),
            "stages": stages, "spectra_methods": methods, "roofline": roofline, "gpu_launches": int(launches), "parity": parity,
            "result_check": {"peak_bin_series0": peak_bin, "peak_freq_THz": float(FREQ[peak_bin]), "checksum": checksum}}
    if e2e:
        line["e2e"] = e2e
    if sampler:
        line["clocks"] = sampler.window(w0, w1)
        sampler.stop()
    if world == 1 and not args.skip_cpu:
        from oracle import port  # noqa: F401  (cpu_baseline leg: the oracle as the CPU baseline)
        vsub = np.concatenate([gen_chunk_torch(torch, dev, wl, t0, t0 + CHUNK).cpu().numpy()
                               for t0 in range(0, args.cpu_frames, CHUNK)])[: args.cpu_frames]
        q_idx = [0, 1237, 2560, 5119][: args.cpu_q]
        eig_sub = eig[q_idx].cpu().numpy()
        res = cpu_sample(wl, vsub, eig_sub, q_idx, args.cpu_series, frames)
        line["cpu_baseline"] = {
            "value": res["value"], "unit": "atom*frames/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
            "sample": f"projection of {args.cpu_frames} frames x {wl['n_atoms']} atoms onto {len(q_idx)} q "
                      f"({res['t_proj_sample_s']:.2f} s) + contraction ({res['t_phonon_sample_s']:.3f} s) + FFT spectra of "
                      f"{args.cpu_series} series of {frames} samples ({res['t_fft_sample_s']:.2f} s); scaled linearly "
                      "to the full job (extrapolated)",
            "full_job_extrapolated_s": res["t_full_extrapolated_s"],
            "note": "oracle port = the reference's numpy projection / np.correlate+np.fft spectrum; this path is serial "
                    "numpy in the reference (no OpenMP), so one host thread does the work"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workload", default="cfg5-A", choices=["cfg5-A", "cfg5-B"],
                    help="cfg5-A (headline): 801 frequencies, 7 windows of 20000; cfg5-B: one window of the whole run, 5244 frequencies")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=131072, help="total frames (default: the full 2^17 workload)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--full-mesh", action="store_true",
                    help="project and store every wave vector instead of one per pair (q, -q) (comparison run)")
    ap.add_argument("--skip-methods", action="store_true")
    ap.add_argument("--sharding", default="auto", choices=["auto", "cyclic", "blocks"],
                    help="N > 1: chunk-cyclic time sharding with overlapped spectra (default) or one time block per rank")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--chunk", type=int, default=4096, help="frames per projection / contraction / exchange chunk")
    ap.add_argument("--cpu-frames", type=int, default=4096, help="frames of the CPU projection sample (BASELINE.md section 3: 4096)")
    ap.add_argument("--write-golden", action="store_true",
                    help="N = 1: write the per-series spectrum sums of this run to gpurun_out/ (committed under tests/golden/ "
                         "as the cross-N parity reference of the multi-GPU runs)")
    ap.add_argument("--ref-procs", type=int, default=0,
                    help="--impl reference: worker processes of the pooled reference sample (0: one per host core; 1: serial)")
    ap.add_argument("--cpu-q", type=int, default=4)
    ap.add_argument("--cpu-series", type=int, default=8)
    args = ap.parse_args()
    if args.workload == "cfg5-B":
        global FREQ, WORKLOAD
        WORKLOAD = "cfg5-B"
        FREQ = np.arange(5244) * (1.0 / (args.frames * DT))
        args.skip_methods = True        # MEM / direct correlation do not depend on the window plan: reported by cfg5-A
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
