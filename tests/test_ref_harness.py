"""The reference arm's driver (baseline/ref_harness.py) -- CPU only, tiny sizes, in a fresh interpreter like bench.py runs it
(the reference's `modules` package and the product's cannot share a process).  Skipped where no copy of the reference exists."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "baseline", "ref_harness.py")


def _have_reference():
    return any(os.path.isfile(os.path.join(p, "src", "dynamic_models2.py")) for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"))


def _run(argv):
    out = subprocess.run([sys.executable, HARNESS] + argv, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert lines, out.stderr[-800:]
    return json.loads(lines[-1])


@pytest.mark.skipif(not _have_reference(), reason="no copy of the reference (baseline/_ref or /root/reference)")
def test_train_step_line():
    r = _run(["--device", "cpu", "--steps", "2", "--warmup", "0", "--batch", "2", "--seq", "10", "20", "20"])
    assert r["impl"] == "reference-unmodified" and r["device"] == "cpu" and r["batch"] == 2
    assert r["ms_per_step"] > 0 and r["ms_per_step_median"] > 0 and r["ms_per_step_max"] >= r["ms_per_step_median"]
    assert abs(r["samples_per_s"] - 2 / (r["ms_per_step"] / 1e3)) < 1e-6 * r["samples_per_s"]


@pytest.mark.skipif(not _have_reference(), reason="no copy of the reference (baseline/_ref or /root/reference)")
def test_ea_fitness_line():
    """the reference's own EvolutionSearch.get_acc over candidates from its own gen_active_cross"""
    r = _run(["--device", "cpu", "--ea", "2", "--valid", "32"])
    assert r["candidates"] == 2 and r["valid_samples"] == 32 and r["subnets_per_s"] > 0
    assert 0.0 <= r["acc_checksum"] <= 2.0            # two binary accuracies
