"""CPU test of bench.py's reference arm (`--impl reference`): the JSON line the driver parses must carry the contract's keys.
The arm times the CPU port of the traversal path (oracle/orc_scene.cpp) -- the one place besides tests/ and smoke() that may run
the oracle -- on a bounded sample, so one step takes well under a second here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("bvh_traversal_microbench")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


# ---- the guard around the extra SPCBPT section (bench.SectionGuard): the headline line must survive a hang or a failure in it ----
_GUARD_SCRIPT = r"""
import json, os, sys, time
sys.path.insert(0, %(root)r)
import bench
rank, mode = int(sys.argv[1]), sys.argv[2]
line = {"metric": "Mrays/s", "value": 1.0} if rank == 0 else None
g = bench.SectionGuard(1, rank, line).start()
if mode == "hang":
    time.sleep(60)                      # a rank stuck in a collective
elif mode == "abandon":
    g.abandon("exception on rank %%d: boom" %% rank)
elif mode == "ok":
    time.sleep(0.2)
    g.finish()
    time.sleep(1.5)                     # the cancelled timer must not print a second line
    if rank == 0:
        line["spcbpt"] = {"samples_per_s": 2.0}
        print(json.dumps(line), flush=True)
    sys.exit(0)
print("not reached")
sys.exit(3)
"""


def _run_guard(rank, mode):
    import time
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-c", _GUARD_SCRIPT % {"root": ROOT}, str(rank), mode], capture_output=True, text=True, timeout=120, cwd=ROOT)
    return r, time.perf_counter() - t0


def test_section_guard_prints_the_headline_when_the_section_hangs():
    r, dt = _run_guard(0, "hang")
    assert r.returncode == 0 and dt < 30, (r.returncode, dt, r.stderr[-1000:])
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and lines[0]["value"] == 1.0 and "timed out" in lines[0]["spcbpt"]["error"]
    assert "still running" in r.stderr and "time.sleep" not in r.stdout       # the stacks go to stderr
    # the other ranks leave quietly (a few seconds after rank 0), status 0
    r, dt = _run_guard(1, "hang")
    assert r.returncode == 0 and dt < 30 and not r.stdout.strip()


def test_section_guard_prints_once_when_the_section_returns():
    r, _ = _run_guard(0, "ok")
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1 and lines[0]["spcbpt"] == {"samples_per_s": 2.0}


def test_section_guard_abandon_keeps_the_headline():
    r, _ = _run_guard(0, "abandon")
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1 and "boom" in lines[0]["spcbpt"]["error"] and "not reached" not in r.stdout
    r, _ = _run_guard(1, "abandon")
    assert r.returncode == 0 and not r.stdout.strip()


def _dry_run(mode):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dry_run.py"), mode], capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    return r, lines


def test_bench_main_flow_on_a_fake_device():
    """bench.main() end to end on a fake CUDA surface (tests/bench_dry_run.py): one JSON line with every contract key, the parity leg
    green, and the SPCBPT section's three outcomes -- result, exception, hang -- all leave the headline line intact"""
    for mode, check in (("ok", lambda sp: sp == {"samples_per_s": 123.0}), ("raise", lambda sp: "boom in section" in sp["error"]),
                        ("hang", lambda sp: "timed out" in sp["error"]), ("norender", lambda sp: sp is None)):
        r, lines = _dry_run(mode)
        assert r.returncode == 0 and len(lines) == 1, (mode, r.returncode, r.stdout[-500:], r.stderr[-1500:])
        d = lines[0]
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                    "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity"):
            assert key in d, (mode, key)
        assert d["metric"] == "Mrays/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0
        assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "hbm"
        assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] == 3 * 48 * 48 * 32
        assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["config"]["workload"]
        p = d["parity"]
        assert p["rays_checked"] == 3 * 48 * 48 and p["prim_mismatch"] == p["tuv_bit_mismatch"] == p["visibility_mismatch"] == 0
        assert check(d.get("spcbpt")), (mode, d.get("spcbpt"))


def test_bench_main_flow_two_ranks_under_torchrun():
    """the same dry run launched as the driver launches N > 1 (torch.distributed.run, 2 ranks, gloo instead of NCCL): barriers and the max over
    ranks run; a rank that hangs or fails inside the section neither takes the headline line down nor makes the launcher fail"""
    import socket
    for mode, check in (("ok2", lambda sp: sp == {"samples_per_s": 123.0}), ("hang1", lambda sp: "timed out" in sp["error"]),
                        ("raise1", lambda sp: "error" in sp)):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                            os.path.join(ROOT, "tests", "bench_dry_run.py"), mode], capture_output=True, text=True, timeout=400, cwd=ROOT)
        lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
        assert r.returncode == 0 and len(lines) == 1, (mode, r.returncode, r.stdout[-500:], r.stderr[-1500:])
        assert lines[0]["n_gpus"] == 2 and lines[0]["value"] > 0 and lines[0]["parity"]["prim_mismatch"] == 0 and check(lines[0]["spcbpt"]), (mode, lines[0].get("spcbpt"))
