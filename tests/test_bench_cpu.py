"""bench.py's contract where no GPU is needed: the reference arm (the reference's own render() compiled from its sources,
timed on the host cores) prints ONE JSON line with the keys the driver reads, on the same `config` object the repo arm
reports; ranks other than 0 of a multi-rank launch do no work; the repo arm fails loudly without a GPU instead of
falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)


def _run(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env)


def _have_ref(w, h):
    from oracle.ref_oracle import have
    return have(w, h)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--config", "1", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames_per_s" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    import bench
    from malevich_b200 import scenes
    assert d["config"] == bench.workload_config(1, scenes.CONFIGS[1]())  # the repo arm's config object, key for key
    if not _have_ref(1280, 720):
        assert "unavailable" in d
        return
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-6 * 1e3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "full frames" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0", "--gpus", "2"], {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_repo_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--config", "1", "--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-traffic"])
    assert r.returncode != 0
    assert r.stdout.strip() == ""  # no line, no fallback number
