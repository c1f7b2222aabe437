"""B200 solver: the reference's ``sw_solver.numpy`` entry points on sm_100a kernels.

Same names, positional signatures and error behaviour as ``src/sw_solver/numpy.py``:

* ``solve``                 numpy.py:409-575 -- the time loop runs on the GPU (one fused
  kernel per step, dt / time / step count device-resident, no per-step host sync);
* ``lax_wendroff_update``   numpy.py:136-356;
* ``DiffusionCoefficients`` numpy.py:12-63, ``compute_diffusion`` 359-382,
  ``compute_laplacian`` 66-133, ``_apply_bcs`` 385-406;
* ``Stepper``               (new) the device-resident loop body for callers that drive
  their own loop or already hold the state on the GPU.

Python only prepares 1-D tables and moves pointers; all arithmetic of the path happens in
``libswb200.so`` (hand-written CUDA, see csrc/), reached through the C ABI in
include/swb200.h.  PyTorch allocates device memory and provides streams.  There is no CPU
fallback: without a CUDA device every entry point raises.
"""

import ctypes
import math
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DTYPES, MODES, SwbConfig, SwbError, SwbStatus, SwbTables, check
from .grid import CartesianGrid, LatLonGrid
from .ic import ICType, SeparableCoriolis, coriolis_latitude, get_initial_conditions, zonal_flow_coriolis
from .utils import EARTH_CONSTANTS, FloatArray2D, FloatT

#: grids with more cells than this get their initial state assembled on the device (nothing M x N on the host)
_DEVICE_IC_CELLS = 1 << 24

#: save_data snapshots are kept on the device up to this many bytes (and a quarter of the free memory), then on the host
_SNAPSHOT_DEVICE_BUDGET = 16 << 30

#: arithmetic mode used when a caller does not choose: "fast" (FMA contraction, shared
#: reciprocals; <= 1e-12 from numpy) or "strict" (bit-identical to numpy, slower)
DEFAULT_MODE = "fast"


# --------------------------------------------------------------------------------------
# host-side helpers (pure numpy / python; covered by CPU tests)
# --------------------------------------------------------------------------------------
def band_layout(N: int, world: int, halo: int) -> List[Dict[str, int]]:
    """Split the updated latitude rows 1..N-2 into ``world`` contiguous bands.

    Each band holds ``halo`` extra rows beyond each updated edge (the pole bands hold the
    single copy row 0 / N-1 instead).  Returns, per rank, the fields of ``swb_config``:
    j_first, j_last (inclusive, global), j_begin, n_local.
    """
    rows = N - 2
    if world < 1 or rows < world:
        raise ValueError(f"cannot split {rows} latitude rows into {world} bands")
    out = []
    base, extra = divmod(rows, world)
    j = 1
    for r in range(world):
        n = base + (1 if r < extra else 0)
        j_first, j_last = j, j + n - 1
        j_begin = max(0, j_first - halo)
        j_end = min(N, j_last + 1 + halo)
        out.append({"rank": r, "j_first": j_first, "j_last": j_last, "j_begin": j_begin, "n_local": j_end - j_begin})
        j += n
    return out


def _pad_to(vec: np.ndarray, n: int, lead: int = 0) -> np.ndarray:
    """Place ``vec`` at offset ``lead`` of a length-n vector, repeating the end values."""
    out = np.empty(n, dtype=np.float64)
    out[lead:lead + len(vec)] = vec
    out[:lead] = vec[0]
    out[lead + len(vec):] = vec[-1]
    return out


def diffusion_weights_lat(cart_grid: CartesianGrid) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The latitude set Ay, By, Cy of numpy.py:50-63 as length-N vectors (edge rows
    duplicated exactly as the reference pads them)."""
    d0, d1 = cart_grid.dy1d[:-1], cart_grid.dy1d[1:]
    Ay = (d1 - d0) / (d1 * d0)
    By = d0 / (d1 * (d1 + d0))
    Cy = -d1 / (d0 * (d1 + d0))
    N = cart_grid.N
    return _pad_to(Ay, N, 1), _pad_to(By, N, 1), _pad_to(Cy, N, 1)


def latitude_tables(latlon_grid: LatLonGrid, cart_grid: CartesianGrid, f_lat: Optional[np.ndarray],
                    use_diffusion: bool, j_begin: int = 0, n_local: Optional[int] = None,
                    f_sep: Optional[SeparableCoriolis] = None) -> Dict[str, np.ndarray]:
    """The 1-D tables ``swb_set_tables`` takes, sliced to a band (include/swb200.h)."""
    N = latlon_grid.N
    n_local = N if n_local is None else n_local
    sl = slice(j_begin, j_begin + n_local)
    full = {
        "c": latlon_grid.c1d,
        "tg": latlon_grid.tg1d,
        "ac": cart_grid.ac1d,
        "tgMidy": _pad_to(latlon_grid.tgMidy1d, N),
        "cMidy": _pad_to(latlon_grid.cMidy1d, N),
        "dy": _pad_to(cart_grid.dy1d, N),
        "dy1": _pad_to(cart_grid.dy11d, N),
        "dyc": _pad_to(cart_grid.dyc1d, N, 1),
        "dy1c": _pad_to(cart_grid.dy1c1d, N, 1),
    }
    if f_lat is not None:
        full["f_lat"] = np.asarray(f_lat, dtype=np.float64)
    if use_diffusion:
        full["Ay"], full["By"], full["Cy"] = diffusion_weights_lat(cart_grid)
    if f_sep is not None:
        full["fsep_lat_b"], full["fsep_lat_d"] = f_sep.lat_b, f_sep.lat_d
    out = {k: np.ascontiguousarray(v[sl], dtype=np.float64) for k, v in full.items()}
    out["phi"] = np.ascontiguousarray(latlon_grid.phi1d, dtype=np.float64)
    if f_sep is not None:
        out["fsep_lon"] = f_sep.lon
    return out


def _latitude_only(f) -> Optional[np.ndarray]:
    """Return f as a latitude vector if it does not depend on longitude, else None."""
    if isinstance(f, torch.Tensor):
        f = f.detach().cpu().numpy()
    f = np.asarray(f)
    if f.ndim == 1:
        return np.ascontiguousarray(f, dtype=np.float64)
    if f.strides[0] == 0 or bool((f == f[0:1]).all()):
        return np.ascontiguousarray(f[0], dtype=np.float64)
    return None


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("sw_solver_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("sw_solver_b200 runs on CUDA devices only")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


def _to_host(t: torch.Tensor) -> np.ndarray:
    """A device tensor as a fresh numpy array.  The destination is touched first by `torch.zeros` (its fill runs on all host
    threads): a D2H copy into never-touched pageable memory is bound by the page faults of the ONE thread inside the driver's
    memcpy -- 2.2 GB/s against 10.5 GB/s this way (`scripts/probe_d2h_result.py`, 388 MB of stacked snapshots)."""
    if t.numel() * t.element_size() < (1 << 22):
        return t.cpu().numpy()
    host = torch.zeros(t.shape, dtype=t.dtype)
    host.copy_(t)
    return host.numpy()


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# --------------------------------------------------------------------------------------
# the device-resident loop body
# --------------------------------------------------------------------------------------
class Stepper:
    """Device-resident state + the fused step kernel for one grid (or one latitude band).

    ``h, u, v`` (and a 2-D ``f`` / ``hs`` when given) are (M+3, N) arrays in the
    reference layout -- numpy arrays are uploaded, CUDA tensors are used in place -- or,
    with ``layout="device"``, CUDA tensors already in the kernels' latitude-major layout
    (n_local, M+3).  With ``band`` (an entry of ``band_layout``) the stepper owns rows
    ``j_begin .. j_begin+n_local-1`` only and the arrays may be either the full grid or
    that window.  After construction ``enqueue(n)`` runs n iterations of numpy.py:496-549
    without involving the host; ``status()`` synchronises and reports time / step count.
    """

    def __init__(self, latlon_grid: LatLonGrid, cart_grid: CartesianGrid, h, u, v, f, hs=None, *,
                 courant: float, final_time: float, use_diffusion: bool = True, mode: Optional[str] = None,
                 device=None, log_capacity: int = 0, band: Optional[Dict[str, int]] = None, world: int = 1,
                 layout: str = "reference", dtype: str = "float64"):
        mode = DEFAULT_MODE if mode is None else mode
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}, got {mode!r}")
        dtype = {"f64": "float64", "f32": "float32"}.get(str(dtype), str(dtype)).replace("torch.", "")
        if dtype not in DTYPES:
            raise ValueError(f"dtype must be 'float64' or 'float32', got {dtype!r}")
        if dtype == "float32" and mode == "strict":
            raise ValueError("the fp32 mode has no strict (bit-exact float64) arithmetic; use mode='fast'")
        self.dtype = dtype
        self._tdtype = torch.float64 if dtype == "float64" else torch.float32
        if layout not in ("reference", "device"):
            raise ValueError("layout must be 'reference' or 'device'")
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.latlon_grid, self.cart_grid = latlon_grid, cart_grid
        self.M, self.N = latlon_grid.M, latlon_grid.N
        self.mode = mode
        self.use_diffusion = bool(use_diffusion)
        M, N = self.M, self.N
        if band is None:
            band = {"rank": 0, "j_first": 1, "j_last": N - 2, "j_begin": 0, "n_local": N}
            world = 1
        self.band, self.world = dict(band), int(world)
        self.j_begin, self.n_local = int(band["j_begin"]), int(band["n_local"])
        self.halo = (2 if self.use_diffusion else 1) if self.world > 1 else 0
        self._layout = layout
        self._peer_bases: List[int] = []

        # Device layout: latitude-major (n_local, pitch), longitude contiguous (include/swb200.h)
        self.pitch = int(self.lib.swb_required_pitch(M, DTYPES[dtype]))
        if self.pitch <= 0:
            raise ValueError(f"bad grid size M={M}")
        self._raw = torch.zeros((2, 3, self.n_local, self.pitch), dtype=self._tdtype, device=self.device)
        #: (2, 3, M+3, n_local) view of the two buffer sets in the reference's index order
        self.buf = self._raw[..., : M + 3].transpose(-1, -2)

        if isinstance(hs, torch.Tensor):
            hs_flat = not bool((hs != 0).any())  # flat topography only subtracts an exact zero
        else:
            hs_flat = hs is None or not np.any(np.asarray(hs))
        # Coriolis: a latitude vector (length N, or the band's window), a separable parameter (tables), or a
        # genuinely 2-D field
        f_sep: Optional[SeparableCoriolis] = None
        if isinstance(f, SeparableCoriolis):
            f_sep, f_lat = f, None
        elif isinstance(f, (np.ndarray, torch.Tensor)) and f.ndim == 1:
            f_lat = np.ascontiguousarray(f.detach().cpu().numpy() if isinstance(f, torch.Tensor) else f, dtype=np.float64)
        else:
            f_lat = _latitude_only(f) if layout == "reference" else None
            if f_lat is None and layout == "reference" and isinstance(f, np.ndarray) and os.environ.get("SWB_FSEP", "1") != "0":
                # the reference's zonal-flow Coriolis array (ic.py:126-133), recognised bit for bit: rebuilt in-kernel
                cand = zonal_flow_coriolis(latlon_grid)
                if cand.matches(f):
                    f_sep = cand
        if f_sep is not None and (not hs_flat or dtype != "float64"):
            f, f_sep = f_sep.array(), None  # topography / fp32: the array kernels (fp32 then refuses, as for any 2-D f)
        if f_lat is not None and len(f_lat) == self.n_local != N:  # the band's window: place it
            full = np.zeros(N)
            full[self.j_begin:self.j_begin + self.n_local] = f_lat
            f_lat = full

        cfg = SwbConfig()
        cfg.struct_bytes = ctypes.sizeof(SwbConfig)
        cfg.M, cfg.N = M, N
        cfg.j_begin, cfg.n_local = self.j_begin, self.n_local
        cfg.j_first, cfg.j_last, cfg.pitch = int(band["j_first"]), int(band["j_last"]), self.pitch
        cfg.dtype, cfg.mode = DTYPES[dtype], MODES[mode]
        cfg.use_diffusion = int(self.use_diffusion)
        cfg.f_is_2d = int(f_lat is None and f_sep is None)
        cfg.f_sep = int(f_sep is not None)
        if f_sep is not None:
            cfg.fsep_scale, cfg.fsep_sa = f_sep.scale, f_sep.sa
        cfg.has_hs = int(not hs_flat)
        cfg.device = self.device.index
        cfg.log_capacity = int(log_capacity)
        cfg.rank, cfg.world, cfg.halo = int(band.get("rank", 0)), self.world, self.halo
        cfg.courant, cfg.final_time = float(courant), float(final_time)
        cfg.dxmin, cfg.dymin = float(cart_grid.dxmin), float(cart_grid.dymin)
        cfg.g, cfg.a, cfg.nu = EARTH_CONSTANTS.g, EARTH_CONSTANTS.a, EARTH_CONSTANTS.nu
        cfg.dphi = float(latlon_grid.dphi)
        self.cfg = cfg
        # host arrays: the copy of the first field is under way while the context is created (tables, tensor maps ...)
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._stage_pair: List[Optional[torch.Tensor]] = [None, None]
        self._prefetched: Optional[Tuple[torch.Tensor, torch.cuda.Event]] = None
        if layout == "reference":
            self._prefetch(h)
        self._ctx = ctypes.c_void_p()
        try:
            check(self.lib.swb_create(ctypes.byref(self._ctx), ctypes.byref(cfg)))
        except Exception:
            if self._copy_stream is not None:  # the prefetched copy must not outlive its staging buffer
                self._copy_stream.synchronize()
            raise

        tabs = latitude_tables(latlon_grid, cart_grid, f_lat, self.use_diffusion, self.j_begin, self.n_local, f_sep)
        ct = SwbTables()
        for name in _lib.TABLE_FIELDS:
            arr = tabs.get(name)
            setattr(ct, name, arr.ctypes.data_as(_lib._dp) if arr is not None else None)
        check(self.lib.swb_set_tables(self._ctx, ctypes.byref(ct)))

        # set 0 receives the initial state (H2D copies and K4 transposes pipelined over two streams)
        self._upload((h, u, v), [self._raw[0, k] for k in range(3)])
        self._f2d = self._upload((f,), [None])[0] if (f_lat is None and f_sep is None) else None
        self._hs = None if hs_flat else self._upload((hs,), [None])[0]

        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        check(self.lib.swb_bind_state(self._ctx, *(ptr(self._raw[s, k]) for s in (0, 1) for k in range(3)),
                                      ptr(self._f2d), ptr(self._hs)))
        self._cfl_ready = False
        self._staging: Optional[torch.Tensor] = None

    # -- plumbing ------------------------------------------------------------------
    def _side_stream(self) -> torch.cuda.Stream:
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        return self._copy_stream

    def _host_window(self, q) -> torch.Tensor:
        """A field given in the reference layout as the (M+3, n_local) float64 tensor this stepper owns (validated)."""
        M, nl = self.M, self.n_local
        t = q if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64))
        if t.dtype != torch.float64 or t.ndim != 2 or t.shape[0] != M + 3 or t.shape[1] not in (self.N, nl):
            raise ValueError(f"expected a float64 array of shape {(M + 3, self.N)} (or the band window {(M + 3, nl)}), "
                             f"got {tuple(t.shape)} {t.dtype}")
        if t.shape[1] != nl:  # the full grid was given: take this band's rows
            t = t[:, self.j_begin:self.j_begin + nl]
        return t

    def _prefetch(self, q) -> None:
        """Start the H2D copy of a PINNED host field into the first staging buffer (copy stream); `_upload` picks it up."""
        if not (isinstance(q, torch.Tensor) and not q.is_cuda and q.is_pinned()):
            return
        t = self._host_window(q)
        main, side = torch.cuda.current_stream(self.device), self._side_stream()
        self._stage_pair[0] = torch.empty((self.M + 3, self.n_local), dtype=torch.float64, device=self.device)
        side.wait_stream(main)  # the allocation is ordered on the current stream
        with torch.cuda.stream(side):
            self._stage_pair[0].copy_(t, non_blocking=True)
            landed = torch.cuda.Event()
            landed.record(side)
        self._prefetched = (t, landed)

    def _upload(self, fields, dsts: List[Optional[torch.Tensor]]) -> List[torch.Tensor]:
        """Bring fields into the device layout (``dsts[k]`` or a fresh (n_local, pitch) array each).

        Host arrays: the H2D copy of field k+1 (copy stream, two staging buffers) overlaps the K4
        transpose of field k (current stream); pinned sources are copied asynchronously.  CUDA
        tensors in the reference layout are transposed in place of the copy; ``layout="device"``
        tensors are copied (and rounded to the state type in the fp32 mode)."""
        M, nl = self.M, self.n_local
        out: List[torch.Tensor] = []
        main = torch.cuda.current_stream(self.device)
        staging = self._stage_pair  # two (M+3, n_local) device buffers, kept for the download of the result
        freed: List[Optional[torch.cuda.Event]] = [None, None]  # staging buffer k is free again (its transpose ran)
        nhost = 0
        for q, dst in zip(fields, dsts):
            if dst is None:
                dst = torch.zeros((nl, self.pitch), dtype=self._tdtype, device=self.device)
            out.append(dst)
            if self._layout == "device":
                if not (isinstance(q, torch.Tensor) and q.is_cuda and tuple(q.shape) == (nl, M + 3) and q.dtype == torch.float64):
                    raise ValueError(f"layout='device' expects float64 CUDA tensors of shape {(nl, M + 3)}")
                dst[:, : M + 3].copy_(q)  # rounds to the state type in the fp32 mode
                continue
            t = self._host_window(q)
            if t.is_cuda:
                src = t if t.stride(1) == 1 else t.contiguous()
                check(self.lib.swb_to_device_layout(self._ctx, ctypes.c_void_p(src.data_ptr()), int(src.stride(0)),
                                                    ctypes.c_void_p(dst.data_ptr()), _stream(self.device)))
                continue
            k = nhost & 1
            nhost += 1
            side = self._side_stream()
            pre, self._prefetched = self._prefetched, None
            if pre is not None and k == 0 and pre[0].data_ptr() == t.data_ptr() and pre[0].shape == t.shape and pre[0].stride() == t.stride():
                landed = pre[1]  # this field's copy into staging[0] was started by `_prefetch`
            else:
                if pre is not None:
                    side.wait_event(pre[1])  # (a prefetched field that is not the one asked for: its staging buffer is reused)
                if staging[k] is None:
                    staging[k] = torch.empty((M + 3, nl), dtype=torch.float64, device=self.device)
                    side.wait_stream(main)  # the allocation is ordered on the current stream
                if freed[k] is not None:
                    side.wait_event(freed[k])
                with torch.cuda.stream(side):
                    staging[k].copy_(t, non_blocking=True)  # H2D: plumbing
                    landed = torch.cuda.Event()
                    landed.record(side)
            main.wait_event(landed)
            check(self.lib.swb_to_device_layout(self._ctx, ctypes.c_void_p(staging[k].data_ptr()), int(staging[k].stride(0)),
                                                ctypes.c_void_p(dst.data_ptr()), _stream(self.device)))
            freed[k] = torch.cuda.Event()
            freed[k].record(main)
        if nhost:
            main.synchronize()  # the sources may be temporaries
        return out

    def release_staging(self):
        """Give the two (M+3, n_local) staging buffers of ``_upload`` / ``download`` back to the allocator."""
        torch.cuda.current_stream(self.device).synchronize()
        self._stage_pair = [None, None]

    def _load(self, q, dst: Optional[torch.Tensor]) -> torch.Tensor:
        return self._upload((q,), [dst])[0]

    def download(self, current: int, out) -> None:
        """K4 + D2H of the buffer set ``current`` into three host tensors of shape (M+3, n_local)
        (pinned ones are filled asynchronously): the transpose of field k+1 (current stream) overlaps
        the copy of field k (copy stream), through the same two staging buffers the upload used.
        Returns when all three have landed."""
        M, nl = self.M, self.n_local
        main, side = torch.cuda.current_stream(self.device), self._side_stream()
        copied: List[Optional[torch.cuda.Event]] = [None, None]
        for k in range(3):
            b = k & 1
            if self._stage_pair[b] is None:
                self._stage_pair[b] = torch.empty((M + 3, nl), dtype=torch.float64, device=self.device)
            if copied[b] is not None:
                main.wait_event(copied[b])  # the copy out of this buffer has finished
            self.export(self._raw[current, k], self._stage_pair[b])
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                out[k].copy_(self._stage_pair[b], non_blocking=True)
                copied[b] = torch.cuda.Event()
                copied[b].record(side)
        side.synchronize()

    def export(self, src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device-layout (n_local, pitch) array -> (M+3, n_local) CUDA tensor in the reference layout."""
        if out is None:
            out = torch.empty((self.M + 3, self.n_local), dtype=torch.float64, device=self.device)
        check(self.lib.swb_from_device_layout(self._ctx, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                              int(out.stride(0)), _stream(self.device)))
        return out

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            torch.cuda.synchronize(self.device)
            for base in self._peer_bases:
                self.lib.swb_ipc_close(ctypes.c_void_p(base))
            self._peer_bases = []
            self.lib.swb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # -- the path ------------------------------------------------------------------
    def cfl_init(self):
        """numpy.py:503-521 on the full initial state; time and step count restart at 0."""
        check(self.lib.swb_cfl_init(self._ctx, _stream(self.device)))
        self._cfl_ready = True

    def enqueue(self, n: int):
        """Enqueue n loop iterations (numpy.py:496-549).  Returns immediately."""
        if not self._cfl_ready:
            self.cfl_init()
        check(self.lib.swb_enqueue_steps(self._ctx, int(n), _stream(self.device)))

    def step_fixed_dt(self, dt: float):
        """One update with a caller-given dt, buffer set 0 -> set 1 (numpy.py:136-356, 541-549)."""
        check(self.lib.swb_step_fixed_dt(self._ctx, float(dt), _stream(self.device)))

    def status_async(self, pinned: Optional[torch.Tensor] = None):
        """Enqueue a status snapshot into ``pinned`` (uint8, >= sizeof(swb_status))."""
        dst = ctypes.c_void_p(pinned.data_ptr()) if pinned is not None else None
        check(self.lib.swb_status_async(self._ctx, dst, _stream(self.device)))

    def status(self) -> SwbStatus:
        st = SwbStatus()
        check(self.lib.swb_get_status(self._ctx, ctypes.byref(st), _stream(self.device)))
        return st

    def state(self, current: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Device views (M+3, n_local) of the current h, u, v (synchronises unless ``current`` is given)."""
        cur = self.status().current if current is None else current
        return self.buf[cur, 0], self.buf[cur, 1], self.buf[cur, 2]

    def snapshot(self, current: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """K4: the buffer set ``current`` as one (3, M+3, n_local) CUDA tensor in the reference
        layout (numpy.py:488-492, 559-563), produced on the current stream."""
        if out is None:
            out = torch.empty((3, self.M + 3, self.n_local), dtype=torch.float64, device=self.device)
        for k in range(3):
            self.export(self._raw[current, k], out[k])
        return out

    def to_numpy(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        snap = _to_host(self.snapshot(self.status().current))
        return snap[0], snap[1], snap[2]

    def max_speed(self) -> float:
        """max sqrt(u^2+v^2) over the full current state (numpy.py:553-554)."""
        out = ctypes.c_double()
        check(self.lib.swb_max_speed(self._ctx, ctypes.byref(out), _stream(self.device)))
        return out.value

    def max_speed_async(self, current: int, out: torch.Tensor):
        """Enqueue K5 on buffer set ``current``; the result (a float64) lands in ``out`` -- 8 bytes of
        pinned host or device memory -- in stream order.  No synchronisation."""
        check(self.lib.swb_max_speed_async(self._ctx, int(current), ctypes.c_void_p(out.data_ptr()), _stream(self.device)))

    def read_log(self, count: int) -> Tuple[np.ndarray, np.ndarray]:
        dts, ts = np.empty(count), np.empty(count)
        check(self.lib.swb_read_log(self._ctx, dts.ctypes.data_as(_lib._dp), ts.ctypes.data_as(_lib._dp),
                                    int(count), _stream(self.device)))
        return dts, ts

    def laplacian(self, q) -> torch.Tensor:
        """(D_x D_x + D_y D_y) q at the interior cells (numpy.py:359-382), as (M+1, N-2)."""
        qd = self._load(q, None)
        out = torch.zeros_like(qd)
        check(self.lib.swb_laplacian(self._ctx, ctypes.c_void_p(qd.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                     _stream(self.device)))
        return out[1:-1, 1:self.M + 2].transpose(0, 1)

    def laplacian_ext(self, q_ext) -> torch.Tensor:
        """The reference's ``compute_laplacian`` (numpy.py:66-133) on an ALREADY extended (M+5, N+2)
        array -- whatever extension the caller built is used as it is -- as (M+1, N-2)."""
        M, N = self.M, self.N
        t = q_ext if isinstance(q_ext, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q_ext, dtype=np.float64))
        if t.dtype != torch.float64 or tuple(t.shape) != (M + 5, N + 2):
            raise ValueError(f"expected a float64 array of shape {(M + 5, N + 2)}, got {tuple(t.shape)} {t.dtype}")
        qe = t.to(self.device).t().contiguous()  # (N+2, M+5): latitude-major like the state (plumbing)
        out = torch.zeros((self.n_local, self.pitch), dtype=torch.float64, device=self.device)
        check(self.lib.swb_laplacian_ext(self._ctx, ctypes.c_void_p(qe.data_ptr()), int(qe.stride(0)), ctypes.c_void_p(out.data_ptr()),
                                         _stream(self.device)))
        return out[1:-1, 1:M + 2].transpose(0, 1)

    def reload(self, h, u, v, f=None, hs=None):
        """Replace the state of buffer set 0 (and a 2-D ``f`` / ``hs`` of a context created with them) by
        new arrays of the same kind -- the operator-level entry points reuse one context per grid."""
        self._upload((h, u, v), [self._raw[0, k] for k in range(3)])
        self._cfl_ready = False  # the next `enqueue` restarts the loop (time 0) from the CFL maxima of this state
        if self._f2d is not None and f is not None:
            self._upload((f,), [self._f2d])
        if self._hs is not None and hs is not None:
            self._upload((hs,), [self._hs])

    @property
    def launch_count(self) -> int:
        return int(self.lib.swb_launch_count(self._ctx))

    @property
    def resident(self) -> bool:
        """True when ``enqueue(n)`` is ONE launch of the resident kernel (grids that fit the
        GPU as one wave of blocks) instead of one launch per step."""
        return bool(self.lib.swb_resident(self._ctx))

    @property
    def streaming(self) -> bool:
        """True when the resident kernel runs its streaming variant (grids beyond one wave of blocks:
        one contiguous range of rows per block, row records travelling with the TMA tiles)."""
        return int(self.lib.swb_resident(self._ctx)) == 2

    # -- latitude bands (multi-GPU) ---------------------------------------------------
    def exchange_block(self) -> Tuple[int, int]:
        """Device address and size of the block peers deposit their CFL maxima in."""
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_uint64()
        check(self.lib.swb_exchange_block(self._ctx, ctypes.byref(ptr), ctypes.byref(nbytes)))
        return int(ptr.value), int(nbytes.value)

    def buffer_addresses(self, base: Optional[int] = None) -> List[int]:
        """Addresses of h0,u0,v0,h1,u1,v1 (optionally relative to another mapping's base)."""
        b = self._raw.data_ptr() if base is None else base
        stride = self.n_local * self.pitch * self._raw.element_size()
        return [b + (s * 3 + k) * stride for s in (0, 1) for k in range(3)]

    def link(self, exchange_blocks: List[int], south: Optional[Dict[str, Any]], north: Optional[Dict[str, Any]]):
        """Hand the context its peers (include/swb200.h swb_link_peers).  ``exchange_blocks``:
        every rank's block as mapped in this process; ``south`` / ``north``: dicts with the
        neighbour's ``bufs`` (six mapped addresses), ``j_begin`` and ``pitch``, or None."""
        xb = (ctypes.c_void_p * len(exchange_blocks))(*exchange_blocks)

        def pack(nb):
            if nb is None:
                return None, 0, 0
            return (ctypes.c_void_p * 6)(*nb["bufs"]), int(nb["j_begin"]), int(nb["pitch"])

        sb, sj, sp = pack(south)
        nb_, nj, np_ = pack(north)
        check(self.lib.swb_link_peers(self._ctx, xb, sb, sj, sp, nb_, nj, np_))
        self._cfl_ready = False


def link_local(steppers: List["Stepper"]):
    """Link latitude bands that live in ONE process on ONE device (plain device pointers).
    Serves tests and single-GPU emulation of the band path: launches of the bands are then
    enqueued round-robin on one stream, so every wait in a step prologue is already met."""
    blocks = [s.exchange_block()[0] for s in steppers]
    for r, s in enumerate(steppers):
        nb = lambda t: {"bufs": t.buffer_addresses(), "j_begin": t.j_begin, "pitch": t.pitch}  # noqa: E731
        s.link(blocks, nb(steppers[r - 1]) if r > 0 else None, nb(steppers[r + 1]) if r + 1 < len(steppers) else None)


# --------------------------------------------------------------------------------------
# operator-level drop-ins (numpy.py:12-63, 66-133, 136-356, 359-406)
# --------------------------------------------------------------------------------------
class DiffusionCoefficients:
    """Diffusion weight container with the reference's attributes Ax..Cy (numpy.py:12-63).

    The kernels need only the latitude vectors (``Ay1d`` ...) -- the longitude set is
    rebuilt in registers from ``x = ac_j * phi_i`` -- so the 2-D arrays are materialised
    on first access only.
    """

    def __init__(self, cart_grid: CartesianGrid):
        self.cart_grid = cart_grid
        self.Ay1d, self.By1d, self.Cy1d = diffusion_weights_lat(cart_grid)
        self._x = None

    def _xset(self):
        if self._x is None:
            dx = self.cart_grid.dx
            d0, d1 = dx[:-1, 1:-1], dx[1:, 1:-1]
            wrap = lambda w: np.concatenate((w[-2:-1, :], w, w[1:2, :]), axis=0)  # noqa: E731  numpy.py:35
            self._x = (wrap((d1 - d0) / (d1 * d0)), wrap(d0 / (d1 * (d1 + d0))), wrap(-d1 / (d0 * (d1 + d0))))
        return self._x

    Ax = property(lambda self: self._xset()[0])
    Bx = property(lambda self: self._xset()[1])
    Cx = property(lambda self: self._xset()[2])

    def _yview(self, v):
        return np.broadcast_to(v[None, :], (self.cart_grid.M + 1, self.cart_grid.N))

    Ay = property(lambda self: self._yview(self.Ay1d))
    By = property(lambda self: self._yview(self.By1d))
    Cy = property(lambda self: self._yview(self.Cy1d))


def _grids_of(cart_grid: CartesianGrid) -> LatLonGrid:
    return LatLonGrid(cart_grid.M, cart_grid.N)


def _match(result: torch.Tensor, like):
    """Return device results in the kind of array the caller passed in."""
    if isinstance(like, torch.Tensor):
        return result.clone() if like.is_cuda else result.cpu()
    return _to_host(result)


# One context per (grid, arithmetic mode, device, kind of Coriolis / topography input) is kept alive for
# the operator-level entry points: a caller that drives its own loop -- the reason these exist -- pays
# the context set-up (a dozen allocations, table upload, tensor maps) once, not per call.
_OPERATOR_CONTEXTS: "Dict[tuple, Stepper]" = {}
_OPERATOR_CONTEXT_LIMIT = 6
_OPERATOR_CONTEXT_CELLS = 1 << 24  # larger grids (> 0.8 GB of state per context) are not kept


def _release_operator_contexts():
    while _OPERATOR_CONTEXTS:
        try:
            _OPERATOR_CONTEXTS.popitem()[1].close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass


import atexit  # noqa: E402

atexit.register(_release_operator_contexts)


def _operator_key(latlon_grid, f, hs, use_diffusion, mode, dev):
    f_lat = None
    if f is not None:
        f_lat = np.ascontiguousarray(f, dtype=np.float64) if (isinstance(f, np.ndarray) and f.ndim == 1) else _latitude_only(f)
    if isinstance(hs, torch.Tensor):
        hs_flat = not bool((hs != 0).any())
    else:
        hs_flat = hs is None or not np.any(np.asarray(hs))
    sep = f_lat is None and isinstance(f, np.ndarray) and f.ndim == 2 and hs_flat and os.environ.get("SWB_FSEP", "1") != "0" \
        and zonal_flow_coriolis(latlon_grid).matches(f)
    return (latlon_grid.M, latlon_grid.N, mode, dev.index, bool(use_diffusion), f_lat is None and f is not None, not hs_flat,
            None if f_lat is None else f_lat.tobytes(), bool(sep))


def lax_wendroff_update(latlon_grid: LatLonGrid, cart_grid: CartesianGrid, dt: FloatT, f: FloatArray2D,
                        hs: FloatArray2D, h: FloatArray2D, u: FloatArray2D, v: FloatArray2D, *,
                        mode: Optional[str] = "strict", device=None):
    """One Lax-Wendroff update (numpy.py:136-356) on the GPU.

    Takes (M+3, N) arrays (numpy or torch), returns ``hnew, unew, vnew`` of shape
    (M+1, N-2) in the same kind of array; inputs are not modified
    (reference tests/numpy_test.py:33-80).  Defaults to the bit-exact arithmetic mode.
    The context of the grid is kept for the next call (see ``_OPERATOR_CONTEXTS``).
    """
    mode = DEFAULT_MODE if mode is None else mode
    dev = _require_cuda(device)
    key = _operator_key(latlon_grid, f, hs, False, mode, dev)
    st = _OPERATOR_CONTEXTS.pop(key, None)
    if st is None:
        while len(_OPERATOR_CONTEXTS) >= _OPERATOR_CONTEXT_LIMIT:
            _OPERATOR_CONTEXTS.pop(next(iter(_OPERATOR_CONTEXTS))).close()
        st = Stepper(latlon_grid, cart_grid, h, u, v, f, hs, courant=0.0, final_time=0.0, use_diffusion=False,
                     mode=mode, device=dev)
    else:
        st.reload(h, u, v, f, hs)
    keep = (latlon_grid.M + 3) * latlon_grid.N <= _OPERATOR_CONTEXT_CELLS
    if keep:
        _OPERATOR_CONTEXTS[key] = st  # most recently used last
    try:
        st.step_fixed_dt(dt)
        out = tuple(_match(st.buf[1, k, 1:-1, 1:-1], h) for k in range(3))
        torch.cuda.current_stream(st.device).synchronize()
    finally:
        if not keep:
            st.close()
    return out


def _diffusion_context(diffusion: "DiffusionCoefficients", device) -> Stepper:
    cart = diffusion.cart_grid
    dev = _require_cuda(device)
    key = (cart.M, cart.N, "laplacian", dev.index)
    st = _OPERATOR_CONTEXTS.pop(key, None)
    if st is None:
        while len(_OPERATOR_CONTEXTS) >= _OPERATOR_CONTEXT_LIMIT:
            _OPERATOR_CONTEXTS.pop(next(iter(_OPERATOR_CONTEXTS))).close()
        zeros = torch.zeros((cart.M + 3, cart.N), dtype=torch.float64, device=dev)
        st = Stepper(_grids_of(cart), cart, zeros, zeros, zeros, np.zeros(cart.N), None, courant=0.0, final_time=0.0,
                     use_diffusion=True, mode="strict", device=dev)
    _OPERATOR_CONTEXTS[key] = st
    return st


def compute_diffusion(quantity: FloatArray2D, diffusion: DiffusionCoefficients, *, device=None):
    """Laplacian of a (M+3, N) quantity at the interior cells, shape (M+1, N-2)
    (numpy.py:359-382: periodic extension in longitude, duplicated rows in latitude)."""
    st = _diffusion_context(diffusion, device)
    out = _match(st.laplacian(quantity), quantity)
    torch.cuda.current_stream(st.device).synchronize()
    return out


def compute_laplacian(coeff: DiffusionCoefficients, q: FloatArray2D, *, device=None):
    """Laplacian of an already-extended (M+5, N+2) array (numpy.py:66-133), shape (M+1, N-2).

    Like the reference it applies the difference weights to ``q`` as given -- ``q[2:-2, 2:-2]`` is the
    centre, the outer two rows / columns are whatever the caller put there (``compute_diffusion``
    builds the reference's own extension, numpy.py:376-381)."""
    st = _diffusion_context(coeff, device)
    out = _match(st.laplacian_ext(q), q)
    torch.cuda.current_stream(st.device).synchronize()
    return out


def _apply_bcs(q, qnew):
    """Boundary conditions (numpy.py:385-406): periodic wrap in longitude, zero-gradient
    copy at the pole rows; ``q`` is modified in place and returned.  Inside ``solve`` this
    is fused into the step kernel; this stand-alone form serves callers of the operator API
    (numpy arrays or torch tensors on any device)."""
    cat = torch.cat if isinstance(q, torch.Tensor) else np.concatenate
    q[:, 1:-1] = cat((qnew[-2:-1, :], qnew, qnew[1:2, :]), 0)
    q[:, 0] = q[:, 1]
    q[:, -1] = q[:, -2]
    return q


# --------------------------------------------------------------------------------------
# the driver
# --------------------------------------------------------------------------------------
_STATUS_BYTES = ctypes.sizeof(SwbStatus)
_STATUS_SLOT = 64  # bytes per status slot (a swb_status record and the K5 result behind it)


def _read_status(pinned: torch.Tensor) -> SwbStatus:
    return SwbStatus.from_buffer_copy(pinned.numpy().tobytes()[:_STATUS_BYTES])


class _PinnedArena:
    """Page-locked host memory handed out in slices: ONE ``cudaHostAlloc`` per chunk instead of one per
    snapshot / status record (numpy.py:488-492, 559-563 append to arrays; a pinned allocation per save
    point would cost more than the step it saves).  Chunks grow geometrically, so a run with
    ``save_data["interval"] = 1`` over thousands of steps makes a handful of allocations."""

    def __init__(self, first_bytes: int, max_chunk_bytes: int = 1 << 30):
        self._next = max(int(first_bytes), 1 << 16)
        self._max = max(int(max_chunk_bytes), self._next)
        self._chunk: Optional[torch.Tensor] = None
        self._used = 0
        self.allocations = 0

    def take(self, nbytes: int) -> torch.Tensor:
        """A pinned uint8 tensor of ``nbytes`` bytes (64-byte aligned inside its chunk)."""
        nbytes = int(nbytes)
        need = (nbytes + 63) // 64 * 64
        if self._chunk is None or self._used + need > self._chunk.numel():
            size = max(self._next, need)
            self._chunk = torch.empty(size, dtype=torch.uint8).pin_memory()
            self._used = 0
            self.allocations += 1
            self._next = min(self._max, 2 * size)
        out = self._chunk[self._used:self._used + nbytes]
        self._used += need
        return out

    def take_f64(self, shape) -> torch.Tensor:
        n = int(np.prod(shape))
        return self.take(8 * n).view(torch.float64).view(*shape)


def _dist_world(distributed: Optional[bool]):
    """(dist module, rank, world) when the band path applies, else (None, 0, 1)."""
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if distributed is True and not on:
        raise RuntimeError("distributed=True needs torch.distributed initialised with more than one rank (torchrun)")
    if distributed is False or not on:
        return None, 0, 1
    return dist, dist.get_rank(), dist.get_world_size()


def solve(num_lon_pts: int, num_lat_pts: int, ic_type: ICType, sim_days: float, courant: float,
          use_diffusion: bool = True, save_data: Optional[Dict[str, Any]] = None, print_interval: int = -1, *,
          mode: Optional[str] = None, device=None, batch_steps: int = 512, dtype: str = "float64",
          distributed: Optional[bool] = None) -> Tuple[FloatArray2D, FloatArray2D, FloatArray2D]:
    """Drop-in for ``sw_solver.numpy.solve`` (numpy.py:409-575) on a B200.

    Same positional parameters, same ``save_data`` contract (keys h, u, v, t, phi, theta
    when ``save_data["interval"] > 0``; first snapshot is the initial state), same error
    order (grid ValueError, negative ``sim_days`` ValueError, non-``ICType`` TypeError).
    Returns the full ghosted (M+3, N) float64 arrays h, u, v as numpy arrays.

    Backend-only keywords: ``mode`` ("fast" | "fast_guarded" | "strict"), ``device``,
    ``batch_steps`` (how many steps are enqueued between two non-blocking termination
    checks), ``dtype`` ("float64" | "float32": the optional fp32 state; results are still
    returned as float64 arrays, within 1e-5 of the reference over benchmark-length runs),
    ``distributed`` (None: latitude bands over all ranks whenever ``torch.distributed`` is
    initialised with more than one rank -- one process per GPU, e.g. under torchrun; every
    rank makes the same call; rank 0 receives the full arrays and ``save_data``, the other
    ranks their own band's window (M+3, n_local) and an untouched ``save_data``).
    """
    save_data = save_data if save_data is not None else {}

    latlon_grid = LatLonGrid(num_lon_pts, num_lat_pts)
    cart_grid = CartesianGrid(latlon_grid)

    if sim_days < 0.0:
        raise ValueError(f"Final time {sim_days} must be non-negative.")
    final_time = 24 * 3600 * sim_days

    if not isinstance(ic_type, ICType):
        raise TypeError(
            f"Invalid problem IC: {ic_type}. See code documentation for implemented initial conditions."
        )
    dev = _require_cuda(device)
    dist, rank, world = _dist_world(distributed)
    M, N = latlon_grid.M, latlon_grid.N

    band = None
    if world > 1:
        band = band_layout(N, world, 2 if use_diffusion else 1)[rank]
    j_begin, n_local = (band["j_begin"], band["n_local"]) if band else (0, N)

    layout = "reference"
    if world > 1 or (M + 3) * N > _DEVICE_IC_CELLS:
        # large grids / bands: assemble the state on the device from the 1-D factors
        # (bit-identical to get_initial_conditions; nothing of size M x N on the host)
        from .bands import rossby_haurwitz_band, zonal_flow_band

        if ic_type == ICType.RossbyHaurwitzWave:
            h, u, v, f = rossby_haurwitz_band(latlon_grid, j_begin, n_local, dev)
        else:
            h, u, v, f = zonal_flow_band(latlon_grid, j_begin, n_local, dev)
        layout = "device"
    else:
        h, u, v, f = get_initial_conditions(ic_type, latlon_grid)
        if ic_type == ICType.RossbyHaurwitzWave:
            f = coriolis_latitude(latlon_grid)

    save_interval = int(save_data.get("interval", 0))
    st = Stepper(latlon_grid, cart_grid, h, u, v, f, None, courant=courant, final_time=final_time,
                 use_diffusion=use_diffusion, mode=mode, device=dev, dtype=dtype, band=band, world=world, layout=layout)
    del h, u, v
    try:
        if world > 1:
            from .bands import link_distributed

            link_distributed(st)
        st.cfl_init()
        # Output pipeline: snapshots and status records land in slices of a few
        # large pinned chunks; the loop below never allocates page-locked memory per step.
        snap_bytes = 3 * (M + 1) * n_local * 8
        arena = _PinnedArena(min(max(4 * snap_bytes, 1 << 20), 1 << 28)) if save_interval > 0 else None
        status_pool = _PinnedArena(256 * _STATUS_SLOT, 1 << 20)
        snaps: List[torch.Tensor] = []   # pinned (3, M+1, n_local) host copies, one per save point
        snap_status: List[torch.Tensor] = []
        stage = None  # device staging of one snapshot in the reference layout (K4)

        # One band: snapshots stay ON THE DEVICE while they fit a budget (a quarter of the free memory, at most
        # 16 GB) and are assembled into the reference's (M+1, N, nsave+1) arrays there -- one strided gather
        # kernel and one D2H per field at the end instead of a host-side np.stack (which, on 742 snapshots of the
        # 1-degree grid, took 30 times as long as the time loop).  Beyond the budget, and for bands (rank 0 gathers
        # host copies), every snapshot goes to pinned host memory as it is taken.
        free_bytes = torch.cuda.mem_get_info(dev)[0]
        dev_budget = min(free_bytes // 4, _SNAPSHOT_DEVICE_BUDGET) if world == 1 else 0
        dev_bytes = 0

        def take_snapshot(current: int) -> torch.Tensor:
            nonlocal stage, dev_bytes
            if dev_bytes + snap_bytes <= dev_budget:
                dev_bytes += snap_bytes
                return st.snapshot(current, None)[:, 1:-1, :]  # numpy.py:488-492: h[1:-1, :]
            stage = st.snapshot(current, stage)
            pinned = arena.take_f64((3, M + 1, n_local))
            for k in range(3):
                pinned[k].copy_(stage[k, 1:-1, :], non_blocking=True)
            return pinned

        if save_interval > 0:
            snaps.append(take_snapshot(0))

        def stops():
            """Step counts at which the host must do something, in increasing order."""
            n = 0
            while True:
                nxt = n + max(1, int(batch_steps))
                for iv in (save_interval, print_interval):
                    if iv > 0:
                        nxt = min(nxt, (n // iv + 1) * iv)
                yield nxt
                n = nxt

        # Every rank sees the same time / step count / termination flag (identical dt on all
        # bands), so the control flow below is the same on every rank.
        batch_slots = [status_pool.take(_STATUS_SLOT) for _ in range(4)]  # reused round-robin: at most two are in flight
        ring: List[Tuple[int, torch.Tensor, torch.cuda.Event]] = []
        enq, nbatch = 0, 0
        for target in stops():
            st.enqueue(target - enq)
            enq = target
            if print_interval > 0 and enq % print_interval == 0:
                # status and K5 (max |velocity| of the buffer set that holds step `enq` if the loop
                # is still running) are enqueued together: ONE synchronisation per print
                slot = status_pool.take(_STATUS_SLOT)
                st.status_async(slot)
                st.max_speed_async(enq & 1, slot[_STATUS_BYTES:_STATUS_BYTES + 8])
                torch.cuda.current_stream(dev).synchronize()
                s = _read_status(slot)
                if s.steps == enq:  # the loop was still running at this step
                    speed = float(slot[_STATUS_BYTES:_STATUS_BYTES + 8].view(torch.float64)[0])
                    if dist is not None:  # numpy's max propagates NaN: so does this one
                        t = torch.tensor([float("inf") if speed != speed else speed], dtype=torch.float64, device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        speed = float("nan") if t.item() == float("inf") else t.item()
                    if rank == 0:
                        print(
                            f"Time = {s.time / 3600.0:6.2f} hours (max {int(final_time / 3600.0)}); "
                            f"max(|u|) = {speed:16.16f}"
                        )
            if save_interval > 0 and enq % save_interval == 0:
                # valid only if all `enq` steps ran, in which case the parity is known
                snaps.append(take_snapshot(enq & 1))
                pin = status_pool.take(_STATUS_SLOT)
                st.status_async(pin)
                snap_status.append(pin)
            pin = batch_slots[nbatch % len(batch_slots)]
            nbatch += 1
            st.status_async(pin)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            ring.append((enq, pin, ev))
            # look at the batch before the one just enqueued: keeps one batch in flight
            if len(ring) >= 2:
                _, pin0, ev0 = ring.pop(0)
                ev0.synchronize()
                s0 = _read_status(pin0)
                if s0.done or s0.error:  # (a device-side error ends the loop on the device too)
                    break
        final = st.status()
        if final.error:
            raise SwbError(final.error, "device-side error flag raised during the time loop "
                                        "(SWB_E_DEPTH: use mode='fast_guarded' or 'strict'; SWB_E_TIMEOUT: a peer GPU never answered)")
        hh, uu, vv = st.to_numpy()

        keep: List[torch.Tensor] = []
        tsave: List[float] = []
        if save_interval > 0:
            torch.cuda.synchronize(dev)
            keep, tsave = [snaps[0]], [0.0]
            for k, pin in enumerate(snap_status):
                s = _read_status(pin)
                if s.steps == (k + 1) * save_interval:  # numpy.py:559: taken while the loop ran
                    keep.append(snaps[k + 1])
                    tsave.append(s.time)

        if world > 1:
            hh, uu, vv, keep = _gather_bands(dist, dev, band, world, N, (hh, uu, vv), keep)

        if save_interval > 0 and rank == 0:
            on_dev = [s for s in keep if isinstance(s, torch.Tensor) and s.is_cuda]
            for idx, name in enumerate(("h", "u", "v")):
                parts = []
                if on_dev:  # (M+1, N, k) assembled on the device, one D2H
                    parts.append(_to_host(torch.stack([s[idx] for s in on_dev], dim=2)))
                rest = [np.asarray(s[idx]) for s in keep[len(on_dev):]]
                if rest:
                    parts.append(np.stack(rest, axis=2))
                save_data[name] = parts[0] if len(parts) == 1 else np.concatenate(parts, axis=2)
            save_data["t"] = np.asarray(tsave)
            # real, writable (M+3, N) arrays like the reference's (numpy.py:570-571), also on grids whose
            # LatLonGrid keeps its 2-D attributes as broadcast views
            save_data["phi"] = np.array(latlon_grid.phi)
            save_data["theta"] = np.array(latlon_grid.theta)
    finally:
        st.close()
    return hh, uu, vv


def _gather_bands(dist, dev, band, world, N, state, snaps):
    """Collect the rows every band owns on rank 0: the final (M+3, N) state and the kept
    snapshots.  Other ranks keep their windows.  Plumbing only (NCCL send / recv)."""
    rank = band["rank"]
    lo = 0 if rank == 0 else band["j_first"]
    hi = N if rank == world - 1 else band["j_last"] + 1
    own = slice(lo - band["j_begin"], hi - band["j_begin"])
    items = [np.stack(state)] + [np.asarray(s) for s in snaps]  # (3, rows_i, n_local) each
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    if rank != 0:
        for a in items:
            dist.send(torch.from_numpy(np.ascontiguousarray(a[:, :, own])).to(dev), dst=0)
        return state[0], state[1], state[2], snaps
    out = []
    for a in items:
        full = np.empty(a.shape[:2] + (N,), dtype=np.float64)
        full[:, :, lo:hi] = a[:, :, own]
        for r in range(1, world):
            rlo, rhi = spans[r]
            buf = torch.empty(a.shape[:2] + (rhi - rlo,), dtype=torch.float64, device=dev)
            dist.recv(buf, src=r)
            full[:, :, rlo:rhi] = buf.cpu().numpy()
        out.append(full)
    return out[0][0], out[0][1], out[0][2], out[1:]
