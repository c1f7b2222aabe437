"""Latitude bands: host logic on the CPU (incl. a world-size-2 gloo run) and,
on a GPU, the in-kernel halo / CFL exchange checked by running several bands on ONE device
(launches round-robin on one stream) against the single-band result -- which must be
bit-identical: the maximum is order-independent and every cell's arithmetic is unchanged."""

import os
import socket

import numpy as np
import pytest
import torch

from sw_solver_b200.bands import neighbour_plan
from sw_solver_b200.grid import CartesianGrid, LatLonGrid
from sw_solver_b200.ic import ICType, get_initial_conditions
from sw_solver_b200.solver import band_layout


def _records(N, world, halo, pitch=80):
    return [dict(b, pitch=pitch) for b in band_layout(N, world, halo)]


def test_neighbour_plan_accepts_a_tiling_and_rejects_gaps():
    recs = _records(64, 4, 2)
    for r in range(4):
        plan = neighbour_plan(list(reversed(recs)), r)  # order of arrival must not matter
        assert plan["exchange"] == [0, 1, 2, 3]
        assert plan["south"] == (r - 1 if r else None) and plan["north"] == (r + 1 if r < 3 else None)
    bad = [dict(d) for d in recs]
    bad[2]["j_first"] += 1
    with pytest.raises(ValueError):
        neighbour_plan(bad, 0)
    with pytest.raises(ValueError):
        neighbour_plan(recs[:2] + recs[3:], 0)
    mixed = [dict(d) for d in recs]
    mixed[1]["pitch"] = 96
    with pytest.raises(ValueError):
        neighbour_plan(mixed, 0)


def test_band_layout_halo_rows():
    for N, world, halo in ((64, 2, 1), (64, 3, 2), (8192 * 8, 8, 1), (20, 6, 2)):
        bands = band_layout(N, world, halo)
        assert bands[0]["j_begin"] == 0 and bands[-1]["j_begin"] + bands[-1]["n_local"] == N
        for b in bands:
            assert b["j_first"] - b["j_begin"] >= (1 if b["rank"] == 0 else halo)
            end = b["j_begin"] + b["n_local"]
            assert end - 1 - b["j_last"] >= (1 if b["rank"] == world - 1 else halo)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, N, halo, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = dict(band_layout(N, world, halo)[rank], pitch=80, xh=b"x%d" % rank, xo=rank, sh=b"s%d" % rank, so=16 * rank)
    infos = [None] * world
    dist.all_gather_object(infos, mine)
    plan = neighbour_plan(infos, rank)
    # the max-over-ranks timing reduction bench.py uses
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = (plan["south"] == (rank - 1 if rank else None) and plan["north"] == (rank + 1 if rank < world - 1 else None)
          and [d["xh"] for d in infos] == [b"x%d" % r for r in range(world)] and t.item() == float(world))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_gloo_world2_band_record_exchange(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_gloo_worker, args=(world, _free_port(), 64, 2, str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(world)] == ["1"] * world


# --------------------------------------------------------------------------------------
# GPU: several bands on one device
# --------------------------------------------------------------------------------------
def _run_bands(M, N, ic, diff, mode, world, nsteps, dtype="float64"):
    from sw_solver_b200.solver import Stepper, link_local

    ll = LatLonGrid(M, N)
    cg = CartesianGrid(ll)
    h, u, v, f = get_initial_conditions(ICType(ic), ll)
    kw = dict(courant=0.5, final_time=1e9, use_diffusion=diff, mode=mode, dtype=dtype)
    one = Stepper(ll, cg, h, u, v, f, None, **kw)
    one.enqueue(nsteps)
    ref = one.to_numpy()
    s_ref = one.status()
    one.close()

    halo = 2 if diff else 1
    bands = band_layout(ll.N, world, halo)
    sts = [Stepper(ll, cg, h, u, v, f, None, band=b, world=world, **kw) for b in bands]
    link_local(sts)
    for st in sts:
        st.cfl_init()
    for _ in range(nsteps):
        for st in sts:  # round-robin on one stream: every prologue wait is already satisfied
            st.enqueue(1)
    out = [np.full_like(ref[0], np.nan) for _ in range(3)]
    stats = []
    for b, st in zip(bands, sts):
        s = st.status()
        stats.append(s)
        loc = st.to_numpy()
        lo = 0 if b["rank"] == 0 else b["j_first"]
        hi = ll.N if b["rank"] == world - 1 else b["j_last"] + 1
        for k in range(3):
            out[k][:, lo:hi] = loc[k][:, lo - b["j_begin"]:hi - b["j_begin"]]
        # the halo rows hold the neighbour's values (stored by the neighbour's kernel)
        for k in range(3):
            assert np.array_equal(loc[k], ref[k][:, b["j_begin"]:b["j_begin"] + b["n_local"]]), (b["rank"], k)
        st.close()
    return ref, out, s_ref, stats


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,ic,diff,mode,world,nsteps,dtype", [
    (96, 64, 0, False, "strict", 2, 12, "float64"),
    (96, 64, 0, True, "strict", 3, 12, "float64"),
    (200, 130, 1, True, "fast", 4, 10, "float64"),     # 2-D Coriolis, ragged band sizes
    (360, 180, 0, False, "fast", 8, 25, "float64"),    # the TMA kernel on every band
    (360, 180, 0, False, "fast", 4, 25, "float32"),    # fp32 state, TMA kernel
    (200, 130, 0, True, "fast", 3, 10, "float32"),     # fp32 state, diffusion
    (360, 180, 0, False, "fast+cflf", 5, 25, "float64"),  # CFL filter forced on: per-band thresholds
])
def test_bands_are_bit_identical_to_one_band(M, N, ic, diff, mode, world, nsteps, dtype, monkeypatch):
    if mode.endswith("+cflf"):
        mode = mode.split("+")[0]
        monkeypatch.setenv("SWB_CFL_FILTER", "1")
    ref, out, s_ref, stats = _run_bands(M, N, ic, diff, mode, world, nsteps, dtype)
    for s in stats:
        assert s.steps == s_ref.steps == nsteps and s.time == s_ref.time and s.error == 0
    # a band's status carries the CFL maxima of ITS rows; their maximum is the global one
    assert max(s.eigenx for s in stats) == s_ref.eigenx and max(s.eigeny for s in stats) == s_ref.eigeny
    for k in range(3):
        assert np.array_equal(out[k], ref[k]), k


@pytest.mark.gpu
def test_device_side_initial_state_is_bit_identical():
    """bands.rossby_haurwitz_band (state assembled on the GPU from 1-D factors) against the
    host path ic.get_initial_conditions, which tests/test_host.py pins to the reference."""
    from sw_solver_b200.bands import rossby_haurwitz_band

    for M, N, jb, nl in ((96, 64, 0, 64), (333, 98, 17, 40), (1440, 720, 700, 20)):
        ll = LatLonGrid(M, N)
        h, u, v, _ = get_initial_conditions(ICType.RossbyHaurwitzWave, ll)
        hd, ud, vd, f = rossby_haurwitz_band(ll, jb, nl, "cuda")
        for a, b in ((hd, h), (ud, u), (vd, v)):
            assert np.array_equal(a.cpu().numpy().T, b[:, jb:jb + nl])
        assert f.shape == (ll.N,)
    # the zonal geostrophic flow (reference ic.py:119-154): state on the device, Coriolis parameter as tables
    from sw_solver_b200.bands import zonal_flow_band

    for M, N, jb, nl in ((96, 64, 0, 64), (333, 98, 17, 40), (1440, 720, 700, 20)):
        ll = LatLonGrid(M, N)
        h, u, v, f2d = get_initial_conditions(ICType.ZonalGeostrophicFlow, ll)
        hd, ud, vd, fs = zonal_flow_band(ll, jb, nl, "cuda")
        for a, b in ((hd, h), (ud, u), (vd, v)):
            assert np.array_equal(a.cpu().numpy().T, b[:, jb:jb + nl])
        assert fs.matches(f2d)
