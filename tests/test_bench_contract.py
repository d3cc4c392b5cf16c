"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm prints ONE JSON line
with the agreed keys (it times the UNMODIFIED reference Model from baseline/_ref when that copy exists, else the CPU
oracle port on a bounded sample), and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, extra=()):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample-users", "1", "--ref-users", "1"] + list(extra), capture_output=True, text=True,
                       timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def _have_ref():
    return any(os.path.isfile(os.path.join(r, "inbatch_sasrec_e2e_text", "model", "model.py"))
               for r in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"))


def test_reference_arm_line():
    lines = _run({"RANK": "0", "WORLD_SIZE": "1"}, ["--cpu-port"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training sequences/sec" and d["unit"] == "sequences/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_unmodified_reference():
    """with a copy of the reference available the arm times the reference's own Model class (kind = reference)"""
    import pytest
    if not _have_ref():
        pytest.skip("no copy of the reference on this box (baseline/_ref is made by __graft_entry__.build())")
    lines = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
    assert d["steps"] == 1 and d["cpu_baseline"]["value"] == d["value"]


def test_reference_arm_other_ranks_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
