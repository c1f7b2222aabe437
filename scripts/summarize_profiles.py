"""Regenerates profiles/r<N>_step_kernel_ncu.md, profiles/r<N>_launches.md and profiles/traffic.json
from the artefacts scripts/gpu_prof.sh (round 1) / scripts/r2_final1.sh (round 2) leave in gpurun_out/
(run here, where ncu reads reports without a GPU):

    python scripts/summarize_profiles.py [--round 2]
"""
import collections
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
import sys

ROUND = int(sys.argv[sys.argv.index("--round") + 1]) if "--round" in sys.argv else 1
CMD = "python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-small" + (" --no-fp32 --no-sustained" if ROUND >= 2 else "")
REP = "prof3.ncu-rep" if ROUND == 1 else f"r{ROUND}_prof_headline.ncu-rep"
PLAIN = "plain2.log" if ROUND == 1 else f"r{ROUND}_plain2.log"
LAUNCHES = f"launches_r{ROUND}.csv"

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_tma_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def step_kernel():
    hdr, units, launches = raw_page(os.path.join(OUT, REP))
    unit = dict(zip(hdr, units))
    out = [f"# ncu --set full, fused step kernel (round {ROUND}, final build)", "",
           f"Command (on a B200 via gpurun, `{'scripts/gpu_prof.sh' if ROUND == 1 else 'scripts/r2_final1.sh'}`; the same command exited 0 without ncu",
           "immediately before); regenerate this file with `python scripts/summarize_profiles.py`:", "",
           "    ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 2 \\",
           f"        -o gpurun_out/{REP[:-8]} {CMD}", "",
           "Workload: Rossby-Haurwitz 16384 x 8192 fp64, diffusion off, mode fast.",
           "Algorithmic bytes per launch: 48 B x 134,193,150 updates = 6.441 GB."]
    traffic = None
    body = []
    for k, r in enumerate(launches):
        d = dict(zip(hdr, r))
        rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tr = rd * scale[unit["dram__bytes_read.sum"]] + wr * scale[unit["dram__bytes_write.sum"]]
        traffic = tr if traffic is None else traffic
        body += ["", f"## launch {k}: {d['Kernel Name']}", "", "| metric | value | unit |", "|---|---|---|"]
        for m in WANT:
            if m in d:
                body.append(f"| {m} | {d[m]} | {unit.get(m, '')} |")
        for m in hdr:
            if "issue_stalled" in m and m.endswith("per_issue_active.ratio"):
                try:
                    if float(d[m]) >= 0.1:
                        body.append(f"| {m} | {d[m]} | {unit.get(m, '')} |")
                except ValueError:
                    pass
        body.append(f"| dram traffic (read + write) / algorithmic bytes | {tr / 6441271200.0:.4f} | |")
    plain = json.loads(open(os.path.join(OUT, PLAIN)).read().strip().splitlines()[-1])
    out += [f"Measured DRAM traffic (read + write) of launch 0: {traffic / 1e9:.3f} GB = {traffic / 6441271200.0:.3f} x algorithmic.",
            f"CUDA-event time of the same command without ncu (bench.py, {plain['steps']} steps): {plain['ms_per_step']:.3f} ms/step,",
            f"roofline fraction {plain['roofline']['frac']:.3f} of the measured {plain['roofline']['peak']} GB/s (burst; sustained runs of 100+ steps under",
            "`sw_power_cap` are 4-8 % slower, see DESIGN.md).", "",
            "SASS of this kernel: `UTMALDG.3D` (one TMA box per tile for h, u, v), `SYNCS.ARRIVE.TRANS64` /",
            "`SYNCS.PHASECHK.TRANS64.TRYWAIT` (mbarrier), `LDS.128`, `STG.E.128`, `SHFL`, `DFMA/DADD/DMUL`, `MUFU.RCP64H/RSQ64H`."]
    open(os.path.join(PROF, f"r{ROUND}_step_kernel_ncu.md"), "w").write("\n".join(out + body) + "\n")
    tj = {"16384x8192:fast": traffic,
          "source": f"profiles/r{ROUND}_step_kernel_ncu.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
    # the other captured kernels, when their raw pages are there (fp32: scripts/gpu_prof_f32.sh; diffusion: scripts/r3_prof_diff.sh)
    for key, raw in (("16384x8192:fast:f32", "r2_prof_f32_raw.csv"), ("7200x3600:fast:diffusion", "r3_prof_diff_raw.csv")):
        path = os.path.join(OUT, raw)
        if os.path.exists(path):
            rows = list(csv.reader(open(path)))
            d = dict(zip(rows[0], rows[2]))
            u = dict(zip(rows[0], rows[1]))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tj[key] = sum(float(d[m]) * scale[u[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    json.dump(tj, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(OUT, LAUNCHES))) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[1:] if len(r) == len(hdr) and r[ix["Metric Name"]] == "gpu__time_duration.sum"]
    names, ns = [], []
    for r in data:
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v *= {"ns": 1, "us": 1e3, "usecond": 1e3, "msecond": 1e6, "ms": 1e6, "nsecond": 1, "second": 1e9}.get(u, 1)
        names.append(r[ix["Kernel Name"]].split("(")[0][:100])
        ns.append(v)
    agg = collections.OrderedDict()
    for n, v in zip(names, ns):
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += v
    total = sum(ns)
    march = sum(v for n, v in zip(names, ns) if "march_kernel" in n)
    fix = sum(v for n, v in zip(names, ns) if "cfl_fixup" in n)
    out = [f"# ncu launch list (round {ROUND}, final build)", "",
           f"    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv {CMD}", "",
           "Per time step ONE `march_kernel` launch followed by ONE `cfl_fixup_kernel` launch (the CFL filter's fix-up: every block",
           "reads a mask and returns).  Everything before the first `march_kernel` is set-up (device-side assembly of the initial",
           "state by torch, the row-record tables, K0 CFL reduction); everything after the last is bench.py's finiteness check.",
           f"Share of `march_kernel` inside the stepping region (march + fix-up launches only): {march / (march + fix):.4f}.", "",
           "| kernel | launches | total ns | share of all profiled launches |", "|---|---|---|---|"]
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n}` | {c} | {v:.0f} | {v / total:.3f} |")
    out += ["", "| id | kernel | ns |", "|---|---|---|"]
    for k, (n, v) in enumerate(zip(names, ns)):
        out.append(f"| {k} | `{n}` | {v:.0f} |")
    open(os.path.join(PROF, f"r{ROUND}_launches.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    step_kernel()
    launch_list()
    print("profiles regenerated")
