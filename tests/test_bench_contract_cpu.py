"""bench.py --impl reference (the CPU arm: the unmodified reference of oracle/_ref -- or the oracle port where that is absent -- timed on
the host cores) prints ONE JSON line with the contract keys, oracle/_ref is a byte-identical copy of the reference's packages,
and the mixed-solve fall-back bookkeeping of the sweep engine works on the CPU stand-in kernels."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.set_default_dtype(torch.float64)


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1",
                          "--warmup", "0", "--ref-rows", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "tensor", "network.py")) or os.path.isdir("/root/reference/tensor")
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_port_fallback_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-port", "--workload", "cfg1", "--steps", "1",
                          "--warmup", "0", "--ref-rows", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_oracle_ref_is_the_unmodified_reference():
    """oracle/_ref (recipe: oracle/make_ref.py) holds the reference's tensor/ and models/ byte for byte."""
    import hashlib
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref_dir, "MANIFEST.json")):
        import pytest
        pytest.skip("oracle/_ref not built (run python oracle/make_ref.py where /root/reference is mounted)")
    man = json.load(open(os.path.join(ref_dir, "MANIFEST.json")))["files"]
    assert "tensor/network.py" in man and "models/tensor_train.py" in man
    for rel, digest in man.items():
        assert hashlib.sha256(open(os.path.join(ref_dir, rel), "rb").read()).hexdigest() == digest, rel
        src = os.path.join("/root/reference", rel)
        if os.path.isfile(src):
            assert open(src, "rb").read() == open(os.path.join(ref_dir, rel), "rb").read(), rel


def test_mixed_solve_fallback_bookkeeping(monkeypatch):
    import fake_ops
    import tensornetworksfork_b200 as tnb
    fake_ops.install(monkeypatch)
    rng = np.random.default_rng(0)
    X = np.concatenate([rng.uniform(-1, 1, size=(200, 3)), np.ones((200, 1))], 1)
    y = np.tanh(X[:, :1])
    layer = tnb.TensorTrainLayer(3, 3, 4, output_shape=1, constrict_bond=False, perturb=True, seed=0)
    net = layer.tensor_network
    net.solve_mode = "mixed"
    net.mixed_accept = -1.0                       # never accept: every solve must be redone by the fp64 path
    assert net.accumulating_swipe(torch.tensor(X), torch.tensor(y), tnb.SquareBregFunction(), method="ridge_cholesky", eps=[1.0, 0.25])
    assert net.solve_stats["mixed"] == 0 and net.solve_stats["mixed_fallback"] == net.solve_stats["fp64"] == 5
    assert net._mixed_floor == 2.0
    net2 = tnb.TensorTrainLayer(3, 3, 4, output_shape=1, constrict_bond=False, perturb=True, seed=0).tensor_network
    net2.solve_mode = "fp64"
    assert net2.accumulating_swipe(torch.tensor(X), torch.tensor(y), tnb.SquareBregFunction(), method="ridge_cholesky", eps=[1.0, 0.25])
    for a, b in zip(net.train_nodes, net2.train_nodes):
        assert torch.equal(a.tensor, b.tensor)
