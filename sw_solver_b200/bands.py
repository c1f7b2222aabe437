"""Latitude bands across GPUs: one process per GPU, launched with torchrun.

The reference is a single process on one array (src/sw_solver/numpy.py:464-494); this
module is the plumbing of the decomposition SURVEY.md section 8(e) describes.  The globe is
cut into ``world`` contiguous latitude bands (``solver.band_layout``).  Every rank builds a
``Stepper`` for its band, then ``link_distributed`` exchanges CUDA IPC handles through
``torch.distributed`` (any backend: the handles are tiny host objects) and maps

* every rank's *exchange block* -- where that rank's step kernel's epilogue deposits the
  CFL maxima of its band, stamped with the step number, and
* the two neighbours' state buffers -- into which the step kernel stores the rows the
  neighbour reads as halo (peer stores over NVLink),

into this process.  From then on the time loop involves no host call and no NCCL
collective: the prologue of step n+1 waits (on its own GPU's memory) until all bands'
maxima of step n have arrived, which also orders the halo rows.  ``torch.distributed`` is
used for set-up, barriers and the timing reduction only.
"""

import ctypes
import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .grid import LatLonGrid
from .ic import separable_rossby_haurwitz, separable_zonal_flow, zonal_flow_coriolis


# --------------------------------------------------------------------------------------
# pure host logic (covered by CPU / gloo tests)
# --------------------------------------------------------------------------------------
def neighbour_plan(infos: List[Dict[str, Any]], rank: int) -> Dict[str, Any]:
    """Which peers' memory rank ``rank`` has to map, from the all-gathered band records.

    Each record carries ``rank, j_begin, n_local, j_first, j_last, pitch`` (+ the opaque
    handles).  Returns ``{"exchange": [ranks in order], "south": rank-1 | None,
    "north": rank+1 | None}`` after checking that the bands tile the interior rows.
    """
    world = len(infos)
    order = sorted(infos, key=lambda d: d["rank"])
    if [d["rank"] for d in order] != list(range(world)):
        raise ValueError("band records do not cover ranks 0..world-1 exactly once")
    for a, b in zip(order[:-1], order[1:]):
        if a["j_last"] + 1 != b["j_first"]:
            raise ValueError(f"bands {a['rank']} and {b['rank']} are not contiguous in latitude")
        if b["j_begin"] > a["j_last"] or a["j_begin"] + a["n_local"] <= b["j_first"]:
            raise ValueError(f"bands {a['rank']} and {b['rank']} do not overlap by their halo rows")
    if len({d["pitch"] for d in order}) != 1:
        raise ValueError("all bands must use the same pitch")
    return {
        "exchange": list(range(world)),
        "south": rank - 1 if rank > 0 else None,
        "north": rank + 1 if rank < world - 1 else None,
    }


def band_record(stepper) -> Dict[str, Any]:
    """The record a rank contributes to the all-gather (IPC handles included)."""
    lib = stepper.lib

    def export(ptr: int) -> Tuple[bytes, int]:
        handle = ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        off = ctypes.c_uint64()
        check(lib.swb_ipc_export(ctypes.c_void_p(ptr), handle, ctypes.byref(off)))
        return handle.raw, int(off.value)

    xptr, _ = stepper.exchange_block()
    xh, xo = export(xptr)
    sh, so = export(stepper._raw.data_ptr())
    b = stepper.band
    return {"rank": int(b["rank"]), "j_begin": stepper.j_begin, "n_local": stepper.n_local,
            "j_first": int(b["j_first"]), "j_last": int(b["j_last"]), "pitch": stepper.pitch,
            "xh": xh, "xo": xo, "sh": sh, "so": so}


# --------------------------------------------------------------------------------------
# linking (needs CUDA)
# --------------------------------------------------------------------------------------
def link_distributed(stepper, group=None) -> None:
    """All-gather the band records over ``torch.distributed`` and link ``stepper``."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world != stepper.world or rank != stepper.band["rank"]:
        raise ValueError("the stepper's band does not match this process's rank / world size")
    infos: List[Optional[Dict[str, Any]]] = [None] * world
    dist.all_gather_object(infos, band_record(stepper), group=group)
    plan = neighbour_plan(infos, rank)
    lib, dev = stepper.lib, stepper.device.index

    def open_handle(raw: bytes) -> int:
        base = ctypes.c_void_p()
        check(lib.swb_ipc_open(raw, dev, ctypes.byref(base)))
        stepper._peer_bases.append(int(base.value))
        return int(base.value)

    blocks = []
    for r in plan["exchange"]:
        if r == rank:
            blocks.append(stepper.exchange_block()[0])
        else:
            blocks.append(open_handle(infos[r]["xh"]) + infos[r]["xo"])

    def neighbour(r: Optional[int]):
        if r is None:
            return None
        d = infos[r]
        base = open_handle(d["sh"]) + d["so"]
        stride = d["n_local"] * d["pitch"] * stepper._raw.element_size()
        return {"bufs": [base + k * stride for k in range(6)], "j_begin": d["j_begin"], "pitch": d["pitch"]}

    stepper.link(blocks, neighbour(plan["south"]), neighbour(plan["north"]))
    # One process per GPU and no two bands on the same device: the bands may run the resident
    # kernel, exchanging their CFL maxima inside the launch (include/swb200.h).
    props = torch.cuda.get_device_properties(stepper.device)
    ident = str(getattr(props, "uuid", "")) or f"{os.uname().nodename}:{dev}"
    idents: List[Optional[str]] = [None] * world
    dist.all_gather_object(idents, ident, group=group)
    if len(set(idents)) == world and os.environ.get("SWB_RESIDENT", "1") != "0":
        check(lib.swb_enable_resident_bands(stepper._ctx))
    dist.barrier(group=group)  # nobody steps before everybody is linked


# --------------------------------------------------------------------------------------
# band-sized initial state, assembled on the device
# --------------------------------------------------------------------------------------
def rossby_haurwitz_band(latlon_grid: LatLonGrid, j_begin: int, n_local: int, device) -> Tuple[torch.Tensor, ...]:
    """h, u, v of the Rossby-Haurwitz wave (reference ic.py:53-107) for latitudes
    ``j_begin .. j_begin+n_local-1``, as (n_local, M+3) float64 CUDA tensors in the device
    layout, plus the Coriolis latitude vector of the WHOLE grid (numpy).

    The transcendental factors are evaluated by numpy on O(M + N) points exactly as
    ``ic.get_initial_conditions`` does; the device only forms the final products and sums,
    one IEEE operation per torch call in the reference's order, so the result is
    bit-identical to slicing the full-grid initial condition.
    """
    from .ic import coriolis_latitude

    s = separable_rossby_haurwitz(latlon_grid)
    sl = slice(j_begin, j_begin + n_local)
    dev = torch.device(device)

    def col(name):  # latitude factor -> column vector
        return torch.from_numpy(np.ascontiguousarray(s[name][sl])).to(dev)[:, None]

    def row(name):  # longitude factor -> row vector
        return torch.from_numpy(np.ascontiguousarray(s[name])).to(dev)[None, :]

    cR, c2R, sR = row("cosR"), row("cos2R"), row("sinR")
    t = col("hB") * cR
    t = col("hA") + t
    t2 = col("hC") * c2R
    t = t + t2
    del t2
    # a DEVICE divisor: torch turns `tensor / python_scalar` into a multiplication by the
    # reciprocal, which is not the reference's IEEE division
    t = t / torch.tensor(float(s["g"]), dtype=torch.float64, device=dev)
    h = float(s["h0"]) + t
    del t
    u = col("uB") * cR
    u = col("uA") + u
    v = col("vB") * sR
    return h, u, v, coriolis_latitude(latlon_grid)


def zonal_flow_band(latlon_grid: LatLonGrid, j_begin: int, n_local: int, device) -> Tuple[Any, ...]:
    """h, u, v of the zonal geostrophic flow (reference ic.py:119-154) for latitudes
    ``j_begin .. j_begin+n_local-1`` as (n_local, M+3) float64 CUDA tensors in the device layout, plus
    the Coriolis parameter of the WHOLE grid in separable form (``ic.SeparableCoriolis``: no 2-D
    array anywhere).  As in ``rossby_haurwitz_band`` numpy evaluates the transcendental factors on
    O(M + N) points and the device forms the final products and sums one IEEE operation per torch
    call in the reference's order: bit-identical to slicing ``get_initial_conditions``."""
    s = separable_zonal_flow(latlon_grid)
    sl = slice(j_begin, j_begin + n_local)
    dev = torch.device(device)
    M3 = latlon_grid.M + 3

    def col(name):  # latitude factor -> column vector
        return torch.from_numpy(np.ascontiguousarray(s[name][sl])).to(dev)[:, None]

    def row(name):  # longitude factor -> row vector
        return torch.from_numpy(np.ascontiguousarray(s[name])).to(dev)[None, :]

    def scalar(name):  # a DEVICE scalar: torch folds python-scalar divisors into reciprocals
        return torch.tensor(float(s[name]), dtype=torch.float64, device=dev)

    sa = scalar("sa")
    rot = col("ct") * row("ncp")      # (-cos(phi) * cos(theta)): the product commutes exactly
    rot = rot * sa
    rot = rot + col("sca")
    t = rot * rot                      # numpy evaluates ** 2.0 as a square
    del rot
    t = scalar("K") * t
    t = t / scalar("g")
    h = scalar("h0") - t
    del t
    u = col("st") * row("cp")
    u = u * sa
    u = col("cca") + u
    u = scalar("u0") * u
    v = (row("sp_nu0") * sa).expand(n_local, M3).contiguous()
    return h, u, v, zonal_flow_coriolis(latlon_grid)


# --------------------------------------------------------------------------------------
# self-check of the decomposition (G bands must be BIT-IDENTICAL to one band)
# --------------------------------------------------------------------------------------
def parity_check(M: int, N: int, use_diffusion: bool, mode: str, nsteps: int, device, group=None) -> Optional[Dict[str, Any]]:
    """Run under torchrun (one rank per GPU).  Every rank steps its latitude band of the M x N
    Rossby-Haurwitz globe for ``nsteps`` steps over the real peer mappings (``link_distributed``);
    rank 0 also steps the WHOLE globe on its own GPU and compares each band's window -- halo, ghost
    and pole rows included -- bit for bit, together with step count, time and error flag.  All
    comparisons happen on the device.  Returns the record on rank 0, None elsewhere."""
    import torch.distributed as dist

    from .grid import CartesianGrid
    from .ic import coriolis_latitude
    from .solver import Stepper, band_layout

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device(device)
    ll = LatLonGrid(M, N)
    cg = CartesianGrid(ll)
    f_lat = coriolis_latitude(ll)
    kw = dict(courant=0.5, final_time=1e12, use_diffusion=use_diffusion, mode=mode, device=dev, layout="device")
    band = band_layout(ll.N, world, 2 if use_diffusion else 1)[rank]
    hd, ud, vd, _ = rossby_haurwitz_band(ll, band["j_begin"], band["n_local"], dev)
    st = Stepper(ll, cg, hd, ud, vd, f_lat, None, band=band, world=world, **kw)
    del hd, ud, vd
    link_distributed(st, group)
    st.enqueue(nsteps)
    s = st.status()
    resident = st.resident
    mine = st._raw[s.current, :, :, : M + 3].contiguous()  # (3, n_local, M+3), device layout
    st.close()
    del st
    metas: List[Any] = [None] * world
    dist.all_gather_object(metas, (band["j_begin"], band["n_local"], int(s.steps), float(s.time), int(s.error)), group=group)
    rec = None
    if rank == 0:
        hf, uf, vf, _ = rossby_haurwitz_band(ll, 0, ll.N, dev)
        one = Stepper(ll, cg, hf, uf, vf, f_lat, None, **kw)
        del hf, uf, vf
        one.enqueue(nsteps)
        s1 = one.status()
        full = one._raw[s1.current]  # (3, N, pitch)
        ok, bad = True, []
        for r in range(world):
            jb, nl, steps_r, time_r, err_r = metas[r]
            part = mine
            if r > 0:
                part = torch.empty((3, nl, M + 3), dtype=mine.dtype, device=dev)
                dist.recv(part, src=r, group=group)
            same = bool(torch.equal(part, full[:, jb:jb + nl, : M + 3])) and steps_r == s1.steps == nsteps and time_r == s1.time and err_r == 0 == s1.error
            if not same:
                bad.append(r)
            ok = ok and same
            del part
        rec = {"grid": f"{M}x{ll.N}", "diffusion": bool(use_diffusion), "mode": mode, "steps": int(nsteps), "bands": world,
               "bit_identical": bool(ok), "time": float(s1.time), "kernel": "resident" if resident else "launch-per-step"}
        if bad:
            rec["mismatching_ranks"] = bad
        one.close()
        del one, full
    else:
        dist.send(mine, dst=0, group=group)
    del mine
    torch.cuda.empty_cache()
    dist.barrier(group=group)
    return rec
