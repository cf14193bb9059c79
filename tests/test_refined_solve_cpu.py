"""Host logic of the exact refinement in the tensor-core Gram modes (network.py::_solve_refined) on the CPU stand-in kernels:
the Gram only preconditions conjugate gradients on the fp64 matrix-free operator, the fall-back to the fp64 Gram is loud and
remembered, and the sample-sharded path (gloo, world size 2) leaves the iteration collectively."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fake_ops
import golden_util as gu
import test_host_sweep_cpu as hs

torch.set_default_dtype(torch.float64)


def _data(fx):
    x = [torch.tensor(t) for t in fx["x"]] if isinstance(fx["x"], list) else torch.tensor(fx["x"])
    return x, torch.tensor(fx["y"])


@pytest.mark.parametrize("name", ["tt_poly_reg", "tnml_poly_xe", "cpd_reg", "tnml_sincos_qr"])
def test_refined_solve_follows_the_reference_recording(name, monkeypatch):
    """gram_mode='tf32x3' on the stand-ins: every site goes through cholesky_factor + cg on the matrix-free operator and still
    reproduces the reference's per-update losses and final prediction."""
    fake_ops.install(monkeypatch)
    fx = gu.load(name)
    layer = hs.build(fx)
    tn = layer.tensor_network
    tn.gram_mode = "tf32x3"
    tn.small_site_fp64 = 0          # these fixtures are tiny: keep them on the refined path they are here to exercise
    x, y = _data(fx)
    ok, trace = hs.run(fx, layer, x, y)
    assert ok == fx["ok"]
    for (_, _, l), u in zip(trace, fx["updates"]):
        assert abs(l - u["loss"]) <= 1e-7 * max(1.0, abs(u["loss"]))
    assert tn.solve_stats["refined"] == len(trace) and tn.solve_stats["gram_fp64_fallback"] == 0
    pred = tn.forward_batch(x, hs.meta_bs(fx)).numpy()
    assert gu.relerr(pred.reshape(fx["pred"].shape), fx["pred"]) < 1e-7


def test_a_coarse_gram_is_only_a_preconditioner(monkeypatch):
    """Perturb the Gram by 1e-3 (far coarser than TF32): the refined step still solves the fp64 system, in more iterations."""
    fake_ops.install(monkeypatch)
    from tensornetworksfork_b200 import ops
    exact_gram = ops.gram

    def coarse(mode, fa, fb, fc, w, rows, M=None, accumulate=False, flush_rows=None):
        out = exact_gram(mode, fa, fb, fc, w, rows, M=M, accumulate=accumulate)
        if mode != ops.GRAM_FP64:
            g = torch.Generator().manual_seed(int(out.numel()))
            out.mul_(1.0 + 1e-3 * torch.randn(out.shape, generator=g))
        return out

    monkeypatch.setattr(ops, "gram", coarse)
    fx = gu.load("tt_poly_reg")
    x, y = _data(fx)
    ref = hs.build(fx)
    hs.run(fx, ref, x, y)
    layer = hs.build(fx)
    tn = layer.tensor_network
    tn.gram_mode = "tf32x3"
    tn.small_site_fp64 = 0          # these fixtures are tiny: keep them on the refined path they are here to exercise
    ok, trace = hs.run(fx, layer, x, y)
    assert ok and tn.solve_stats["gram_fp64_fallback"] == 0
    assert tn.solve_stats["refine_iters"] > 0          # the perturbed factor is not exact, so the iteration had work to do
    for a, b in zip(ref.tensor_network.train_nodes, tn.train_nodes):
        assert float((a.tensor - b.tensor).norm() / a.tensor.norm()) < 1e-7


def test_refinement_failure_falls_back_to_the_fp64_gram_and_remembers(monkeypatch):
    fake_ops.install(monkeypatch)
    fx = gu.load("tt_poly_reg")
    x, y = _data(fx)
    ref = hs.build(fx)
    hs.run(fx, ref, x, y)
    layer = hs.build(fx)
    tn = layer.tensor_network
    tn.gram_mode = "tf32x3"
    tn.small_site_fp64 = 0          # these fixtures are tiny: keep them on the refined path they are here to exercise
    tn.refine_accept = -1.0             # never accept: every site must be redone with the fp64 Gram
    ok, trace = hs.run(fx, layer, x, y)
    assert ok
    assert tn.solve_stats["refined"] == 0 and tn.solve_stats["gram_fp64_fallback"] >= 1
    assert tn._refine_floor > 0
    # once a ridge has failed, sites with the same or a smaller ridge skip the tensor-core attempt altogether
    assert tn.solve_stats["gram_fp64_fallback"] < len(trace)
    for a, b in zip(ref.tensor_network.train_nodes, tn.train_nodes):
        assert torch.equal(a.tensor, b.tensor)


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_ops as fo
    fo.install()
    fx = gu.load(name)
    layer = hs.build(fx)
    tn = layer.tensor_network
    tn.gram_mode = "tf32x3"
    tn.small_site_fp64 = 0          # these fixtures are tiny: keep them on the refined path they are here to exercise
    N = fx["y"].shape[0]
    cut = [0, N // 2 + 7, N][rank:rank + 2]
    sl = slice(cut[0], cut[1])
    x = [torch.tensor(t[sl]) for t in fx["x"]] if isinstance(fx["x"], list) else torch.tensor(fx["x"][sl])
    y = torch.tensor(fx["y"][sl])
    tn.process_group = dist.group.WORLD
    tn.shard_offset = cut[0]
    tn.shard_total = N
    ok, trace = hs.run(fx, layer, x, y)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ok=ok, losses=np.array([l for _, _, l in trace]), refined=tn.solve_stats["refined"],
             **{f"core{i}": n.tensor.numpy() for i, n in enumerate(tn.train_nodes)})
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tt_poly_reg", "tnml_poly_xe"])
def test_refined_solve_sample_sharded_world2_gloo(name, tmp_path):
    port = hs._free_port()
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    fx = gu.load(name)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert bool(r0["ok"]) and bool(r1["ok"]) and int(r0["refined"]) == len(fx["updates"])
    for i in range(len(fx["cores0"])):
        assert np.array_equal(r0[f"core{i}"], r1[f"core{i}"]), "ranks diverged"
    for l0, l1, u in zip(r0["losses"], r1["losses"], fx["updates"]):
        assert l0 == l1 and abs(l0 - u["loss"]) <= 1e-7 * max(1.0, abs(u["loss"]))
