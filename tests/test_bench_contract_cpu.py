"""CPU: the JSON contract of `bench.py --impl reference` (the CPU restatement timed on the host cores) -- the
arm the driver runs beside the GPU arm -- on a tiny sample; and that ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra):
    env = dict(os.environ, SNRF_REF_RAYS="64", **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train rays/s (fwd+bwd)" and d["unit"] == "rays/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["config"]["workload"] == "default.yaml-single-tile"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_do_no_work():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
