"""bench.py's output contract, as far as it can be exercised without a GPU: the reference arm (the reference's own GATLayer
from oracle/_ref -- or the torch port when that copy is absent -- on a bounded sample) prints ONE JSON line with the agreed
keys and states the N / E' / scale it actually timed, and the B200 arm refuses to run without a CUDA device instead of
falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=900)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--scale", "0.125")   # 1/8 of the default sample: seconds
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gat_layer_fwd_bwd_edges_per_s" and d["unit"] == "edges/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["value"] > 0
    assert "workload" in d["config"] and "products" in d["config"]["workload"]
    # the line describes the bounded sample that was timed, not the B200 arm's full-size graph
    cfg = d["config"]
    assert cfg["bounded_sample"] is True and abs(cfg["scale"] - 0.125 / 64) < 1e-12
    assert 4000 < cfg["n_nodes"] < 6000 and 100_000 < cfg["n_edges_rewritten"] < 150_000
    assert abs(d["value"] - 3 * cfg["n_edges_rewritten"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "models", "gat_layer.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert f"N={cfg['n_nodes']}" in cb["sample"] and f"E'={cfg['n_edges_rewritten']}" in cb["sample"]
    assert isinstance(d["loss"], float) and d["loss"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without CUDA")
def test_b200_arm_has_no_cpu_fallback():
    r = _bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
