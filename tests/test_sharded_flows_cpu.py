"""Sample sharding for every engine, world size 2 over gloo on the CPU stand-in kernels.

Each rank runs the SAME parity case as the single-process tests (the case bodies assert against the reference recordings), but a
patched ``TensorNetwork._prepare_data`` hands the sweep only this rank's rows and switches the engine to its sharded mode
(``process_group``, ``shard_offset``, ``shard_total``): partial Gram / right-hand side / matvec results are all-reduced, the
mean-of-batch-means loss is assembled from global minibatch boundaries.  Uneven shards, and minibatches that straddle the cut."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

torch.set_default_dtype(torch.float64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _install_sharding(rank, world):
    from tensornetworksfork_b200.tensor.network import TensorNetwork, MappedInput
    orig = TensorNetwork._prepare_data

    def rows_of(x):
        if isinstance(x, MappedInput):
            return x.X.shape[0]
        return (x[0] if isinstance(x, (list, tuple)) else x).shape[0]

    def cut_rows(x, lo, hi):
        if isinstance(x, MappedInput):
            return x[lo:hi]
        if isinstance(x, (list, tuple)):
            return [t[lo:hi].contiguous() for t in x]
        return x[lo:hi].contiguous()

    def sharded(self, x, y_true, data_device, model_device):
        N = rows_of(x)
        bounds = [0, N // 2 + 7, N] if N > 20 else [0, N // 2, N]
        lo, hi = bounds[rank], bounds[rank + 1]
        self.process_group, self.shard_offset, self.shard_total = dist.group.WORLD, lo, N
        return orig(self, cut_rows(x, lo, hi), y_true[lo:hi].contiguous(), data_device, model_device)

    TensorNetwork._prepare_data = sharded


def _worker(rank, world, port, which):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_ops
    fake_ops.install()
    _install_sharding(rank, world)
    try:
        if which == "cumsum":
            import cumsum_case
            cumsum_case.run("cpu", teacher_forced=False)
        elif which == "type1":
            import type1_case
            type1_case.run("cpu")
        elif which.startswith("linear:"):
            import linear_cases as lc
            out = lc.run_case(which.split(":", 1)[1], "cpu")
            assert max(float(v) for v in out) < 1e-6, out
        elif which.startswith("krylov:"):
            import krylov_cases as kc
            core_err, loss_err = kc.run_case(which.split(":", 1)[1], "cpu")
            assert core_err < 1e-7 and loss_err < 1e-8, (core_err, loss_err)
        elif which.startswith("conv:"):
            import conv_cases as cc
            fwd, core, loss, pred = cc.run_case(which.split(":", 1)[1], "cpu")
            assert fwd < 1e-12 and core < 1e-8 and loss < 1e-9 and pred < 1e-8, (fwd, core, loss, pred)
        elif which == "conv_grow":
            import conv_cases as cc
            for fwd, core, loss, pred in cc.run_grow("cpu"):
                assert fwd < 1e-12 and core < 1e-8 and loss < 1e-9 and pred < 1e-8, (fwd, core, loss, pred)
        elif which == "dmrg":
            import dmrg_case
            dmrg_case.run("cpu", True)
        elif which.startswith("growing:"):
            import growing_case as gc
            hist_err, pred_err, core_err, score_err = gc.run(which.split(":", 1)[1], "cpu")
            assert hist_err < 1e-8 and pred_err < 1e-7 and core_err < 1e-6, (hist_err, pred_err, core_err)
        else:
            raise ValueError(which)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("which", ["cumsum", "type1", "linear:linear_tt_reg", "linear:linear_tt_xe", "krylov:krylov_lanczos_xe",
                                   "krylov:krylov_cumsum_lanczos", "conv:conv_dense_xe", "conv:conv_type1", "conv_grow", "dmrg",
                                   "growing:a"])
def test_sharded_flow_world2_gloo(which):
    mp.spawn(_worker, args=(2, _free_port(), which), nprocs=2, join=True)
