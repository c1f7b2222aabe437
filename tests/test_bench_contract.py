"""CPU-only checks of bench.py's reference arm (the one leg that runs without a GPU): one JSON
line with the contract's keys on rank 0, silence and exit code 0 on the other ranks."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _run(env_extra, *args):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "360x180", "--steps", "2",
                           "--warmup", "1", *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


def test_reference_arm_prints_one_contract_line():
    res = _run({})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "updates/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "cpu_sample" in d["config"]
    # both arms print the SAME config object (the driver compares the two lines): the function our arm calls, same arguments
    import bench

    assert d["config"] == bench.workload_config(360, 180, 1, False, "f64", False, "fast")
    # the unmodified numpy reference, when its staged copy exists (oracle/_ref, __graft_entry__.build())
    from oracle import ref_numpy

    nr = d["cpu_baseline"]["numpy_ref"]
    if ref_numpy.available():
        assert nr["kind"] == "_ref" and nr["cores"] == 1 and len(nr["runs"]) >= 1
        assert all(r["value"] > 0 and r["seconds_per_step"] > 0 for r in nr["runs"])
    else:
        assert nr is None


def test_reference_arm_under_torchrun_runs_on_rank_zero_only():
    """N > 1: rank 0 prints the line (with every host thread, although torchrun pins
    OMP_NUM_THREADS=1 for its workers); the other ranks exit 0 without output."""
    env = {"WORLD_SIZE": "2", "OMP_NUM_THREADS": "1", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29599"}
    r0 = _run(dict(env, RANK="0", LOCAL_RANK="0"), "--gpus", "2")
    r1 = _run(dict(env, RANK="1", LOCAL_RANK="1"), "--gpus", "2")
    assert r0.returncode == 0 and r1.returncode == 0, (r0.stderr[-1000:], r1.stderr[-1000:])
    assert r1.stdout.strip() == ""
    d = json.loads(r0.stdout.strip().splitlines()[-1])
    assert d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    import bench

    assert d["config"] == bench.workload_config(360, 180, 2, False, "f64", False, "fast")
    assert "360x360 globe in 2 latitude bands" in d["config"]["workload"]
