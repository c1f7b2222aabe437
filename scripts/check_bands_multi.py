"""Multi-process parity check of the latitude-band path (run under torchrun, one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/check_bands_multi.py

Every rank steps its band (real CUDA IPC peer mappings, in-kernel halo stores, exchange-block
CFL maxima); rank 0 also runs the whole grid on its own GPU and compares: the banded result
must be BIT-IDENTICAL.  Covers BASELINE.json configs[2] (1440x720, diffusion
on) and configs[3] (7200x3600, diffusion off).  Prints one line per case; exit code 1 on a
mismatch."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sw_solver_b200 import bands  # noqa: E402
from sw_solver_b200.grid import CartesianGrid, LatLonGrid  # noqa: E402
from sw_solver_b200.ic import coriolis_latitude, rossby_haurwitz_window  # noqa: E402
from sw_solver_b200.solver import Stepper, band_layout  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok_all = True
    for (M, N, diff, mode, nsteps) in ((1440, 720, True, "fast", 40), (1440, 720, True, "strict", 10),
                                       (7200, 3600, False, "fast", 30)):
        ll = LatLonGrid(M, N)
        cg = CartesianGrid(ll)
        f_lat = coriolis_latitude(ll)
        kw = dict(courant=0.5, final_time=1e9, use_diffusion=diff, mode=mode, device=dev)
        band = band_layout(ll.N, world, 2 if diff else 1)[rank]
        h, u, v = rossby_haurwitz_window(ll, band["j_begin"], band["n_local"])
        st = Stepper(ll, cg, h, u, v, f_lat, None, band=band, world=world, **kw)
        bands.link_distributed(st)
        st.enqueue(nsteps)
        s = st.status()
        mine = torch.from_numpy(np.stack(st.to_numpy())).to(dev)  # (3, M+3, n_local)
        st.close()
        # gather the bands' windows on rank 0
        shapes = [None] * world
        dist.all_gather_object(shapes, (band["j_begin"], band["n_local"], band["j_first"], band["j_last"], s.steps, s.time, s.error))
        if rank == 0:
            parts = [mine] + [torch.empty((3, M + 3, shapes[r][1]), dtype=torch.float64, device=dev) for r in range(1, world)]
            for r in range(1, world):
                dist.recv(parts[r], src=r)
            hf, uf, vf = rossby_haurwitz_window(ll, 0, ll.N)
            one = Stepper(ll, cg, hf, uf, vf, f_lat, None, **kw)
            one.enqueue(nsteps)
            s1 = one.status()
            ref = np.stack(one.to_numpy())
            one.close()
            ok = True
            for r in range(world):
                jb, nl = shapes[r][0], shapes[r][1]
                ok = ok and np.array_equal(parts[r].cpu().numpy(), ref[:, :, jb:jb + nl])  # halo rows included
                ok = ok and shapes[r][4] == s1.steps == nsteps and shapes[r][5] == s1.time and shapes[r][6] == 0
            print(f"bands x{world} {M}x{N} diffusion={diff} mode={mode} steps={nsteps}: "
                  f"{'BIT-IDENTICAL to one GPU' if ok else 'MISMATCH'} (t = {s1.time:.6f} s)", flush=True)
            ok_all = ok_all and ok
        else:
            dist.send(mine, dst=0)
        dist.barrier()
    # the driver itself: solve() under torchrun == solve() on one GPU, save_data included
    from sw_solver_b200.ic import ICType
    from sw_solver_b200.solver import solve

    for (M, N, ic, days, diff, mode, interval) in ((96, 64, ICType.RossbyHaurwitzWave, 0.02, True, "strict", 5),
                                                    (200, 130, ICType.ZonalGeostrophicFlow, 0.01, False, "fast", 7),
                                                    (1440, 720, ICType.RossbyHaurwitzWave, 0.002, True, "fast", 10)):
        sd = {"interval": interval}
        h, u, v = solve(M, N, ic, days, 0.5, diff, save_data=sd, mode=mode)
        if rank == 0:
            sd1 = {"interval": interval}
            h1, u1, v1 = solve(M, N, ic, days, 0.5, diff, save_data=sd1, mode=mode, distributed=False)
            ok = all(np.array_equal(a, b) for a, b in ((h, h1), (u, u1), (v, v1)))
            ok = ok and sorted(sd) == sorted(sd1) and all(np.array_equal(sd[k], sd1[k]) for k in ("h", "u", "v", "t", "phi", "theta"))
            print(f"solve() x{world} {M}x{N} {ic.name} diffusion={diff} mode={mode}: "
                  f"{'BIT-IDENTICAL to one GPU' if ok else 'MISMATCH'} ({len(sd1['t'])} snapshots, t_end = {sd1['t'][-1]:.3f} s)", flush=True)
            ok_all = ok_all and ok
        else:
            assert sd == {"interval": interval} and h.shape[0] == M + 3
        dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
