#!/usr/bin/env python
"""Reads the per-rank dumps of a -DSWB_BAND_TIMING build (SWB_BAND_TIMING_FILE=<prefix>) and splits the
step period of the launch-per-step band path into its parts, per rank, over the last `--last` steps:

    wait   gather enter -> exit      waiting for the other bands' stamps (straggler skew + stamp latency)
    gap1   gather exit  -> first block of the step kernel starts          (launch latency)
    march  first block starts -> last block ends                          (the step kernel itself)
    gap2   last block ends -> fix-up kernel starts                        (drain + launch latency)
    gap3   fix-up starts -> publish enters                                (fix-up kernel + launch latency)
    pub    publish enters -> stamp released (st.release.sys behind the outstanding halo stores)
    gap4   stamp released -> next gather enters                           (launch latency)

`last peer` = how often each rank's stamp was the LAST to arrive at this rank (who is the straggler);
`own->last` = time between this rank's own stamp release and the arrival of the last stamp.
Clocks are per-GPU globaltimers (ns): only differences on one GPU are formed.

    python scripts/band_timing.py gpurun_out/bt8.16384x65536 [--last 80]
"""
import argparse
import glob
import sys

import numpy as np


def load(path):
    raw = np.fromfile(path, dtype=np.uint64)
    world, rank, cap, row, seq, resident = (int(x) for x in raw[:6])
    return world, rank, seq, raw[8:].reshape(cap, row).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix")
    ap.add_argument("--last", type=int, default=80)
    ap.add_argument("--skip-tail", type=int, default=4)
    ap.add_argument("--first", type=int, default=0, help="analyse stamps [first, first + last) instead of the last ones (burst regime)")
    a = ap.parse_args()
    files = sorted(glob.glob(a.prefix + ".r*"))
    if not files:
        sys.exit(f"no dumps match {a.prefix}.r*")
    print(f"| rank | period | wait | gap1 | march | gap2 | gap3 | pub | gap4 | own->last | last peer (share of steps) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    rows = []
    for f in files:
        world, rank, seq, t = load(f)
        cap = t.shape[0]
        hi = seq - 1 - a.skip_tail          # last stamp used is seq-1
        lo = hi - a.last
        if a.first > 0:
            lo, hi = a.first, a.first + a.last
        idx = np.arange(lo, hi) % cap
        nxt = np.arange(lo + 1, hi + 1) % cap
        T, Tn = t[idx], t[nxt]
        us = lambda x: float(np.mean(x)) / 1e3  # noqa: E731
        period = us(Tn[:, 0] - T[:, 0]) if world > 1 else us(Tn[:, 2] - T[:, 2])
        if world > 1:
            arr = T[:, 8:8 + world]
            last = np.argmax(arr, axis=1)
            share = np.bincount(last, minlength=world) / len(last)
            prev_pub = t[(np.arange(lo - 1, hi - 1)) % cap][:, 6]
            own_to_last = us(arr.max(axis=1) - prev_pub)
            parts = [us(T[:, 1] - T[:, 0]), us(T[:, 2] - T[:, 1]), us(T[:, 3] - T[:, 2]), us(T[:, 4] - T[:, 3]), us(T[:, 5] - T[:, 4]),
                     us(T[:, 6] - T[:, 5]), us(Tn[:, 0] - T[:, 6])]
            sh = " ".join(f"r{r}:{s:.2f}" for r, s in enumerate(share) if s > 0.005)
        else:
            parts = [0.0, 0.0, us(T[:, 3] - T[:, 2]), us(T[:, 4] - T[:, 3]), 0.0, 0.0, us(Tn[:, 2] - T[:, 4])]
            own_to_last, sh = 0.0, "-"
        rows.append((rank, period, parts))
        print(f"| {rank} | {period:.1f} | " + " | ".join(f"{x:.1f}" for x in parts) + f" | {own_to_last:.1f} | {sh} |")
    print()
    print("(microseconds, means over", a.last, "steps)")


if __name__ == "__main__":
    main()
