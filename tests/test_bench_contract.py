"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm (`--impl reference`: the
reference's own CPU V-cycle through oracle/_ref, or the C port) prints ONE JSON line with the keys the
driver reads, and the thread scan of oracle/cpu_baseline.py reports what it measured."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(built):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-level", "2",
                        "--steps", "2", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "poisson_vcycle_cell_updates_per_s"
    assert d["unit"] == "cell-updates/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["gpu_launches"] == 0 and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["config"]["same_config"] is False and d["config"]["cpu_grid"] == 32  # a bounded sample was asked for


def test_rank_other_than_zero_of_the_reference_arm_does_nothing(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--cpu-level", "2", "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cited_evidence_files_exist():
    """every profiles/ file that DESIGN.md, README.md or profiles/README.md cites by name is committed"""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    have = {os.path.basename(p) for p in glob.glob(os.path.join(root, "profiles", "*"))}
    missing = []
    for doc in ("DESIGN.md", "README.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(root, doc)).read()
        for name in set(re.findall(r"`(?:profiles/)?(r0\d_[A-Za-z0-9_.\-]+\.(?:json|jsonl|csv|txt|tsv))`", text)):
            if name not in have:
                missing.append((doc, name))
    assert not missing, missing
