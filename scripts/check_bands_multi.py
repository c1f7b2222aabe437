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
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
