"""Array-level operators over the C ABI (one method per entry point of include/dynaphopy_b200.h).

``Ops`` is written against a tiny memory back-end so that the SAME marshalling code can be
driven by the test-suite's host emulator (tests/emul.py supplies a numpy back-end together with
the emulator build of the kernels).  The package itself only ever instantiates the CUDA
back-end (``TorchCudaBackend``): device memory, streams and the current device come from
PyTorch, the arithmetic from libdynaphopy_b200.so.  No CPU fallback exists here.
"""
import ctypes

import numpy as np

from . import _lib

UNIT_CONVERSION = 0.00010585723   # u * A^2 * THz -> eV*ps (dynaphopy/power_spectrum/__init__.py:9)


class TorchCudaBackend:
    """Device memory and streams from PyTorch (plumbing only)."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("dynaphopy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)

    def _dtype(self, dtype):
        t = self.torch
        return {np.float64: t.float64, np.complex128: t.complex128, np.int32: t.int32, np.uint8: t.uint8}[dtype]

    def empty(self, shape, dtype):
        return self.torch.empty(tuple(int(s) for s in shape), dtype=self._dtype(dtype), device=self.device)

    def asarray(self, x, dtype):
        t = self.torch
        if isinstance(x, t.Tensor):
            return x.to(device=self.device, dtype=self._dtype(dtype)).contiguous()
        return t.as_tensor(np.ascontiguousarray(x, dtype=dtype), device=self.device)

    def is_complex(self, x):
        return x.is_complex() if isinstance(x, self.torch.Tensor) else np.iscomplexobj(x)

    def ptr(self, x):
        return ctypes.c_void_p(x.data_ptr()) if x is not None else None

    def stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def shape(self, x):
        return tuple(x.shape)


def is_device_array(x):
    return type(x).__module__.startswith("torch")


def to_host(x):
    """Device (torch) array -> numpy, numpy passes through."""
    return x.detach().cpu().numpy() if is_device_array(x) else np.asarray(x)


def shape_of(x):
    return tuple(x.shape)


def _host(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _hptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Ops:
    def __init__(self, lib=None, backend=None):
        self.lib = lib if lib is not None else _lib.load()
        self.be = backend if backend is not None else TorchCudaBackend()

    # ------------------------------------------------------------------ A8 / A9 / A1
    def atomic_displacements(self, trajectory, positions, cell):
        """(T,N,3) positions -> (T,N,3) real displacements (c/displacements.c:95-196 for all atoms)."""
        be = self.be
        cplx = be.is_complex(trajectory)
        traj = be.asarray(trajectory, np.complex128 if cplx else np.float64)
        T, N, _ = be.shape(traj)
        pos = be.asarray(positions, np.float64)
        cellh = _host(cell, np.float64)
        out = be.empty((T, N, 3), np.float64)
        self.lib.call("dp_atomic_displacements", be.ptr(traj), int(cplx), be.ptr(pos), _hptr(cellh), T, N,
                      be.ptr(out), be.stream())
        return out

    def velocity_from_positions(self, trajectory, positions, cell, dt, sqrt_mass=None, prev_frame=None,
                                next_frame=None, want_displacements=False):
        """Fused displacements -> np.gradient -> sqrt(mass) (dynamics.py:158-179, 313-329, 143-156)."""
        be = self.be
        cplx = be.is_complex(trajectory)
        dt_ = np.complex128 if cplx else np.float64
        traj = be.asarray(trajectory, dt_)
        T, N, _ = be.shape(traj)
        pos = be.asarray(positions, np.float64)
        cellh = _host(cell, np.float64)
        sm = be.asarray(sqrt_mass, np.float64) if sqrt_mass is not None else None
        pf = be.asarray(prev_frame, dt_) if prev_frame is not None else None
        nf = be.asarray(next_frame, dt_) if next_frame is not None else None
        out_u = be.empty((T, N, 3), np.float64) if want_displacements else None
        out_v = be.empty((T, N, 3), np.float64)
        self.lib.call("dp_velocity_from_positions", be.ptr(traj), int(cplx), be.ptr(pos), _hptr(cellh), be.ptr(sm),
                      T, N, float(dt), be.ptr(pf), be.ptr(nf), be.ptr(out_u), be.ptr(out_v), be.stream())
        return (out_u, out_v) if want_displacements else out_v

    def displacement_moments(self, trajectory, positions, cell):
        """(T,N,3) positions -> (N,9) per-atom sums over frames of u_k and u_a*u_b (a <= b) of the minimum-image
        displacement (feeds dynamics.py:207-295 without materialising the relative trajectory)."""
        be = self.be
        cplx = be.is_complex(trajectory)
        traj = be.asarray(trajectory, np.complex128 if cplx else np.float64)
        T, N, _ = be.shape(traj)
        pos = be.asarray(positions, np.float64)
        cellh = _host(cell, np.float64)
        sums = be.empty((N, 9), np.float64)
        nbytes = max(self.lib.size("dp_displacement_moments_workspace_bytes", T, N), 256)
        ws = be.empty((nbytes,), np.uint8)
        self.lib.call("dp_displacement_moments", be.ptr(traj), int(cplx), be.ptr(pos), _hptr(cellh), T, N,
                      be.ptr(sums), be.ptr(ws), nbytes, be.stream())
        return sums

    def mass_weight(self, velocity, sqrt_mass):
        be = self.be
        cplx = be.is_complex(velocity)
        v = be.asarray(velocity, np.complex128 if cplx else np.float64)
        T, N, _ = be.shape(v)
        sm = be.asarray(sqrt_mass, np.float64)
        out = be.empty((T, N, 3), np.complex128 if cplx else np.float64)
        self.lib.call("dp_mass_weight", be.ptr(v), 2 if cplx else 1, be.ptr(sm), T, N, be.ptr(out), be.stream())
        return out

    # ------------------------------------------------------------------ A2 / A3
    def project_wave_vector(self, velocity, sqrt_mass, positions, type_idx, n_prim, q_cart, project_on_atom=-1):
        """Dense projection for a batch of Cartesian q: -> (T,Q,n_prim,3) complex128."""
        be = self.be
        cplx = be.is_complex(velocity)
        v = be.asarray(velocity, np.complex128 if cplx else np.float64)
        T, N, _ = be.shape(v)
        q = _host(np.atleast_2d(q_cart), np.float64)
        Q = q.shape[0]
        typ = _host(type_idx, np.int32)
        pos = be.asarray(positions, np.float64)
        sm = be.asarray(sqrt_mass, np.float64) if sqrt_mass is not None else None
        nbytes = self.lib.size("dp_project_wave_vector_workspace_bytes", N, Q)
        ws = be.empty((nbytes,), np.uint8)
        out = be.empty((T, Q, n_prim, 3), np.complex128)
        self.lib.call("dp_project_wave_vector", be.ptr(v), int(cplx), be.ptr(sm), be.ptr(pos), _hptr(typ), N,
                      int(n_prim), T, _hptr(q), Q, int(project_on_atom), be.ptr(out), be.ptr(ws), nbytes, be.stream())
        return out

    def project_commensurate(self, velocity, dims, unit_type, n_prim, qbin, twiddle, project_on_atom=-1, out=None,
                             pair_major=False, validate_bins=True):
        """3-D DFT projection: real (T,N,3) -> (T,Q,n_prim,3) complex128 for commensurate q.
        ``out``: optional preallocated contiguous (T,Q,n_prim,3) destination (e.g. a time slice).
        ``pair_major``: write the intermediate layout (T, 3*n_prim/2, Q, 2) instead (DP_VC_PAIRS), to be
        consumed by ``project_phonon_series(..., pair_major=True)``."""
        be = self.be
        v = be.asarray(velocity, np.float64)
        T, N, _ = be.shape(v)
        dims_h = _host(dims, np.int32)
        ut = _host(unit_type, np.int32)
        n_unit = ut.shape[0]
        assert N == n_unit * int(np.prod(dims_h)), "atom count does not match dims * n_unit"
        if validate_bins and not is_device_array(qbin):
            # host-side bins are checked here; device-resident bins are checked by the kernel itself (a bin outside the
            # grid gives NaN for that q), so the call never waits for the device
            qh = np.asarray(qbin).reshape(-1, 3)
            if qh.size and (qh.min() < 0 or (qh >= dims_h[None, :]).any()):
                raise _lib.DpError(_lib.DP_ERR_INVALID, "project_commensurate: q bins outside the supercell grid")
        qb = be.asarray(qbin, np.int32)
        Q = be.shape(qb)[0]
        tw = be.asarray(twiddle, np.complex128) if twiddle is not None else None      # None: bare lattice DFT
        nbytes = self.lib.size("dp_project_commensurate_workspace_bytes", _hptr(dims_h), n_unit, int(n_prim), Q)
        ws = be.empty((nbytes,), np.uint8)
        shape = (T, 3 * int(n_prim) // 2, Q, 2) if pair_major else (T, Q, int(n_prim), 3)
        if out is None:
            out = be.empty(shape, np.complex128)
        else:
            assert be.shape(out) == shape and out.is_contiguous(), "bad `out` buffer"
        self.lib.call("dp_project_commensurate_ex", be.ptr(v), _hptr(dims_h), n_unit, _hptr(ut), int(n_prim), T,
                      be.ptr(qb), be.ptr(tw), Q, int(project_on_atom), int(bool(pair_major)), be.ptr(out), be.ptr(ws),
                      nbytes, be.stream())
        return out

    def project_commensurate_blocks(self, velocity, dims, unit_type, n_prim, qbin, q_block, block_ptrs=None, out=None):
        """Pair-major bare lattice DFT split into ``Q / q_block`` arrays of shape (T, 3*n_prim/2, q_block, 2), one per
        block of consecutive wave vectors of ``qbin`` (dp_project_commensurate_blocks).  ``block_ptrs``: device address
        of every block (send slots / peer receive buffers); otherwise the blocks are the leading axis of ``out``
        ``(Q / q_block, T, 3*n_prim/2, q_block, 2)`` (allocated when None), which is returned."""
        be = self.be
        v = be.asarray(velocity, np.float64)
        T, N, _ = be.shape(v)
        dims_h = _host(dims, np.int32)
        ut = _host(unit_type, np.int32)
        n_unit = ut.shape[0]
        assert N == n_unit * int(np.prod(dims_h)), "atom count does not match dims * n_unit"
        qb = be.asarray(qbin, np.int32)
        Q = be.shape(qb)[0]
        q_block = int(q_block)
        assert Q % q_block == 0
        nb = Q // q_block
        nbytes = self.lib.size("dp_project_commensurate_workspace_bytes", _hptr(dims_h), n_unit, int(n_prim), Q)
        ws = be.empty((nbytes,), np.uint8)
        if block_ptrs is None:
            shape = (nb, T, 3 * int(n_prim) // 2, q_block, 2)
            if out is None:
                out = be.empty(shape, np.complex128)
            else:
                assert be.shape(out) == shape and out.is_contiguous(), "bad `out` buffer"
            step = T * (3 * int(n_prim) // 2) * q_block * 2 * 16
            base = be.ptr(out).value or be.ptr(ws).value          # (T == 0: nothing is written)
            block_ptrs = [base + g * step for g in range(nb)]
        assert len(block_ptrs) == nb
        ptrs = (ctypes.c_void_p * nb)(*[int(a) for a in block_ptrs])
        self.lib.call("dp_project_commensurate_blocks", be.ptr(v), _hptr(dims_h), n_unit, _hptr(ut), int(n_prim), T,
                      be.ptr(qb), Q, q_block, ptrs, nb, be.ptr(ws), nbytes, be.stream())
        return out

    def project_phonon(self, vc, eigenvectors, q_block=None, out=None, inplace=False):
        """vc (T,Q,n_p,3) or (T,Q,M); eigenvectors (Q,M,n_p,3) or (Q,M,M) -> vq (T,Q,M).

        ``q_block``: write ``(Q/q_block, T, q_block, M)`` instead (all-to-all send layout);
        ``inplace``: overwrite ``vc`` (plain layout only)."""
        be = self.be
        vc = be.asarray(vc, np.complex128)
        T, Q = be.shape(vc)[:2]
        M = int(np.prod(be.shape(vc)[2:]))
        e = be.asarray(eigenvectors, np.complex128)
        assert be.shape(e)[0] == Q and int(np.prod(be.shape(e)[1:])) == M * M, "eigenvector shape mismatch"
        qb = Q if q_block is None else int(q_block)
        if inplace:
            out = vc
        elif out is None:
            out = be.empty((T, Q, M) if qb == Q else (Q // qb, T, qb, M), np.complex128)
        self.lib.call("dp_project_phonon_blocked", be.ptr(vc), be.ptr(e), T, Q, M, qb, be.ptr(out), be.stream())
        return out.reshape(T, Q, M) if (inplace and qb == Q) else out

    def project_phonon_series(self, vc, eigenvectors, q_block=None, out=None, t_offset=0, t_total=None,
                              pair_major=False, block_ptrs=None):
        """vc (T,Q,n_p,3)|(T,Q,M) -> series-major vq: ``out[g, qq*M + s, t_offset + t]`` with q = g*q_block + qq,
        shape (Q/q_block, q_block*M, t_total).  ``out`` lets a long trajectory be contracted chunk by chunk.
        ``block_ptrs``: one device address per block g (possibly peer memory of rank g): block g is written there as a
        (q_block*M, t_total) matrix instead of into ``out`` -- the contraction fused with the exchange step."""
        be = self.be
        vc = be.asarray(vc, np.complex128)
        if pair_major:                                   # (T, M/2, Q, 2) from project_commensurate(pair_major=True)
            T, half_m, Q, _ = be.shape(vc)
            M = 2 * half_m
        else:
            T, Q = be.shape(vc)[:2]
            M = int(np.prod(be.shape(vc)[2:]))
        e = be.asarray(eigenvectors, np.complex128)
        assert be.shape(e)[0] == Q and int(np.prod(be.shape(e)[1:])) == M * M, "eigenvector shape mismatch"
        qb = Q if q_block is None else int(q_block)
        t_total = T if t_total is None else int(t_total)
        if block_ptrs is not None:
            ptrs = (ctypes.c_void_p * len(block_ptrs))(*[int(a) for a in block_ptrs])
            self.lib.call("dp_project_phonon_series_peers", be.ptr(vc), int(bool(pair_major)), be.ptr(e), T, Q, M, qb,
                          int(t_offset), t_total, ptrs, len(block_ptrs), be.stream())
            return None
        if out is None:
            out = be.empty((Q // qb, qb * M, t_total), np.complex128)
        else:
            assert be.shape(out) == (Q // qb, qb * M, t_total), "bad `out` buffer"
        self.lib.call("dp_project_phonon_series_ex", be.ptr(vc), int(bool(pair_major)), be.ptr(e), T, Q, M, qb,
                      int(t_offset), t_total, be.ptr(out), be.stream())
        return out

    def project_phonon_series_mirror(self, vc, eig2, qmap, Q, q_block=None, out=None, t_offset=0, t_total=None,
                                     block_ptrs=None):
        """Half-mesh contraction (dp_project_phonon_series_mirror): ``vc`` (T, M/2, Qh, 2) is the pair-major bare lattice
        DFT of one wave vector of every pair (q, -q); ``qmap`` (Qh, 2) int32 gives the indices of the stored q and of its
        partner in the full list of ``Q`` wave vectors (< 0: none), ``eig2`` (Qh, 2, M, M) the eigenvectors of the first
        and the conjugated eigenvectors of the second.  Output as ``project_phonon_series`` for all Q wave vectors."""
        be = self.be
        vc = be.asarray(vc, np.complex128)
        T, half_m, Qh, _ = be.shape(vc)
        M = 2 * half_m
        e = be.asarray(eig2, np.complex128)
        qm = be.asarray(qmap, np.int32)
        assert be.shape(e) == (Qh, 2, M, M) and be.shape(qm) == (Qh, 2), "mirror tables do not match vc"
        Q = int(Q)
        qb = Q if q_block is None else int(q_block)
        t_total = T if t_total is None else int(t_total)
        if block_ptrs is not None:
            ptrs = (ctypes.c_void_p * len(block_ptrs))(*[int(a) for a in block_ptrs])
            self.lib.call("dp_project_phonon_series_mirror", be.ptr(vc), be.ptr(e), be.ptr(qm), T, Qh, Q, M, qb,
                          int(t_offset), t_total, None, ptrs, len(block_ptrs), be.stream())
            return None
        if out is None:
            out = be.empty((Q // qb, qb * M, t_total), np.complex128)
        else:
            assert be.shape(out) == (Q // qb, qb * M, t_total), "bad `out` buffer"
        self.lib.call("dp_project_phonon_series_mirror", be.ptr(vc), be.ptr(e), be.ptr(qm), T, Qh, Q, M, qb,
                      int(t_offset), t_total, be.ptr(out), None, 0, be.stream())
        return out

    def launch_count(self):
        return int(self.lib.cdll.dp_launch_count())

    # ------------------------------------------------------------------ A7 / A6 / A5
    def _series(self, series):
        s = self.be.asarray(series, np.complex128)
        shp = self.be.shape(s)
        assert len(shp) == 2, "series must be (T, S)"
        return s, shp[0], shp[1]

    def power_fft(self, series, dt, frequency_range, scale=UNIT_CONVERSION, series_major=False):
        """FFT power spectrum of every series -> (F, S) float64.

        ``series``: (T, S) complex128 as the reference passes it, or -- ``series_major=True`` --
        (S, T) with the time axis contiguous (the layout the kernels stream at full bandwidth), or
        (G, S, Tb) = G consecutive time blocks of Tb samples (what the time-sharded ->
        series-sharded all-to-all delivers, T = G*Tb)."""
        be = self.be
        f = _host(frequency_range, np.float64)
        s = be.asarray(series, np.complex128)
        shp = be.shape(s)
        if not series_major:
            assert len(shp) == 2, "series must be (T, S)"
            T, S = shp
            strides = (1, S, 0, 0)
        elif len(shp) == 2:
            S, T = shp
            strides = (T, 1, 0, 0)
        else:
            G, S, Tb = shp
            T = G * Tb
            strides = (Tb, 1, Tb, S * Tb)
        nbytes = self.lib.size("dp_power_fft_workspace_bytes", T, S, float(dt), _hptr(f), f.size)
        if nbytes == 0:
            # let the library explain (bad grid, unsupported window length, ...)
            self.lib.call("dp_power_fft", be.ptr(s), T, S, S, float(dt), _hptr(f), f.size, float(scale), None, None, 0,
                          be.stream())
        ws = be.empty((nbytes,), np.uint8)
        out = be.empty((f.size, S), np.float64)
        self.lib.call("dp_power_fft_strided", be.ptr(s), T, S, strides[0], strides[1], strides[2], strides[3], float(dt),
                      _hptr(f), f.size, float(scale), be.ptr(out), be.ptr(ws), nbytes, be.stream())
        return out

    # -- the same spectrum in steps (series produced incrementally) ----------------------------------------
    def _series_layout(self, series, series_major):
        be = self.be
        s = be.asarray(series, np.complex128)
        shp = be.shape(s)
        if not series_major:
            assert len(shp) == 2, "series must be (T, S)"
            return s, shp[0], shp[1], (1, shp[1], 0, 0)
        if len(shp) == 2:
            return s, shp[1], shp[0], (shp[1], 1, 0, 0)
        g, n, tb = shp
        return s, g * tb, n, (tb, 1, tb, n * tb)

    def power_fft_begin(self, T, S, dt, frequency_range):
        """Workspace of one incremental FFT spectrum (see power_fft_windows / power_fft_finish)."""
        f = _host(frequency_range, np.float64)
        nbytes = self.lib.size("dp_power_fft_workspace_bytes", int(T), int(S), float(dt), _hptr(f), f.size)
        if nbytes == 0:
            self.lib.call("dp_power_fft", None, int(T), int(S), int(S), float(dt), _hptr(f), f.size, 1.0, None, None, 0,
                          self.be.stream())
        return self.be.empty((nbytes,), np.uint8)

    def power_fft_windows(self, series, dt, frequency_range, first_window, n_windows, workspace, series_major=False,
                          max_ctas=0, preemptible=False):
        """Transform windows [first_window, first_window + n_windows) of every series into ``workspace``; reads
        no sample outside those windows (the rest of ``series`` may still be in production).  ``max_ctas``: SMs the
        persistent kernel may occupy (0: all) -- leaves room for concurrent work of other streams.  ``preemptible``:
        one short-lived CTA per item instead (dp_power_fft_windows_preemptible), for callers that rely on stream priorities."""
        be = self.be
        s, T, S, st = self._series_layout(series, series_major)
        f = _host(frequency_range, np.float64)
        if preemptible:
            self.lib.call("dp_power_fft_windows_preemptible", be.ptr(s), T, S, st[0], st[1], st[2], st[3], float(dt), _hptr(f),
                          f.size, int(first_window), int(n_windows), int(max_ctas), be.ptr(workspace),
                          int(workspace.numel() if hasattr(workspace, "numel") else workspace.size), be.stream())
            return
        self.lib.call("dp_power_fft_windows_ex", be.ptr(s), T, S, st[0], st[1], st[2], st[3], float(dt), _hptr(f), f.size,
                      int(first_window), int(n_windows), int(max_ctas), be.ptr(workspace),
                      int(workspace.numel() if hasattr(workspace, "numel") else workspace.size), be.stream())

    def power_fft_finish(self, T, S, dt, frequency_range, workspace, scale=UNIT_CONVERSION):
        be = self.be
        f = _host(frequency_range, np.float64)
        out = be.empty((f.size, S), np.float64)
        self.lib.call("dp_power_fft_finish", int(T), int(S), float(dt), _hptr(f), f.size, float(scale), be.ptr(out),
                      be.ptr(workspace), int(workspace.numel() if hasattr(workspace, "numel") else workspace.size),
                      be.stream())
        return out

    def power_mem(self, series, dt, frequency_range, coefficients, scale=UNIT_CONVERSION, nan_to_num=True,
                  series_major=False):
        """Maximum-entropy spectrum of every series -> (F, S).  ``series`` layouts as in ``power_fft``."""
        be = self.be
        s, T, S, st = self._series_layout(series, series_major)
        f = _host(frequency_range, np.float64)
        nbytes = max(self.lib.size("dp_power_mem_workspace_bytes", T, S, f.size, int(coefficients)), 256)
        ws = be.empty((nbytes,), np.uint8)
        out = be.empty((f.size, S), np.float64)
        self.lib.call("dp_power_mem_strided", be.ptr(s), T, S, st[0], st[1], st[2], st[3], float(dt), _hptr(f), f.size,
                      int(coefficients), float(scale), int(bool(nan_to_num)), be.ptr(out), be.ptr(ws), nbytes, be.stream())
        return out

    def power_correlation(self, series, dt, frequency_range, step=10, integration_method=1,
                          weight=_lib.DP_CORR_WEIGHT_REFERENCE, scale=UNIT_CONVERSION, series_major=False):
        """Direct-correlation spectrum of every series -> (F, S).  ``series`` layouts as in ``power_fft``."""
        be = self.be
        s, T, S, st = self._series_layout(series, series_major)
        f = _host(frequency_range, np.float64)
        nbytes = max(self.lib.size("dp_power_correlation_workspace_bytes", T, S, f.size, int(step)), 256)
        ws = be.empty((nbytes,), np.uint8)
        out = be.empty((f.size, S), np.float64)
        self.lib.call("dp_power_correlation_strided", be.ptr(s), T, S, st[0], st[1], st[2], st[3], float(dt), _hptr(f),
                      f.size, int(step), int(integration_method), int(weight), float(scale), be.ptr(out), be.ptr(ws),
                      nbytes, be.stream())
        return out

    def spectra_peaks(self, spectra, sets=None, want_average=True):
        """(F, S) spectra -> (averaged (F, S) or None, peak_bin (S,), peak_height (S,)).  ``sets``: list of lists of
        series indices (degenerate sets, analysis/fitting/__init__.py:17-39); every series must be in exactly one."""
        be = self.be
        ps = be.asarray(spectra, np.float64)
        F, S = be.shape(ps)
        d_ptr = d_mem = d_sid = None
        n_sets = 0
        if sets is not None:
            ptr = np.zeros(len(sets) + 1, dtype=np.int32)
            ptr[1:] = np.cumsum([len(g) for g in sets])
            mem = np.concatenate([np.asarray(g, dtype=np.int32) for g in sets]) if len(sets) else np.zeros(0, np.int32)
            sid = np.full(S, -1, dtype=np.int32)
            for k, g in enumerate(sets):
                sid[np.asarray(g, dtype=int)] = k
            if mem.size != S or (sid < 0).any() or mem.min() < 0 or mem.max() >= S:
                raise ValueError("degenerate sets must partition the series")
            d_ptr, d_mem, d_sid = be.asarray(ptr, np.int32), be.asarray(mem, np.int32), be.asarray(sid, np.int32)
            n_sets = len(sets)
        avg = be.empty((F, S), np.float64) if (want_average and sets is not None) else None
        bins = be.empty((S,), np.int32)
        heights = be.empty((S,), np.float64)
        self.lib.call("dp_spectra_peaks", be.ptr(ps), F, S, be.ptr(d_ptr), be.ptr(d_mem), be.ptr(d_sid), n_sets,
                      be.ptr(avg), be.ptr(bins), be.ptr(heights), be.stream())
        return avg, bins, heights

    def fft_pieces(self, T, dt, frequency_range, max_pieces=1 << 16):
        """Window list of the FFT method (host query; power_spectrum/__init__.py:33-51)."""
        f = _host(frequency_range, np.float64)
        L = ctypes.c_int(0)
        n = ctypes.c_int(0)
        starts = np.zeros(max_pieces, dtype=np.int32)
        self.lib.call("dp_power_fft_pieces", int(T), float(dt), _hptr(f), f.size, ctypes.byref(L), ctypes.byref(n),
                      _hptr(starts), max_pieces)
        return L.value, [[int(s), int(s) + L.value] for s in starts[:n.value]]

    def fft_engine(self, T, dt, frequency_range):
        """Transform engine that serves this spectrum (dp_power_fft_engine: 1 general, 2 slab, 3 long-window)."""
        f = _host(frequency_range, np.float64)
        rc = self.lib.cdll.dp_power_fft_engine(int(T), float(dt), _hptr(f), f.size)
        if rc < 0:
            raise _lib.DpError(-rc, "dp_power_fft_engine")
        return rc

    def fft_c2c(self, x, direction=-1):
        be = self.be
        x = be.asarray(x, np.complex128)
        batch, n = be.shape(x)
        nbytes = self.lib.size("dp_fft_c2c_workspace_bytes", n, batch)
        out = be.empty((batch, n), np.complex128)
        ws = be.empty((max(nbytes, 256),), np.uint8)
        self.lib.call("dp_fft_c2c", be.ptr(x), be.ptr(out), n, batch, int(direction), be.ptr(ws), nbytes, be.stream())
        return out


_default = None


def default_ops():
    """Process-wide Ops on the current CUDA device."""
    global _default
    if _default is None:
        _default = Ops()
    return _default
