"""Run under torchrun on N GPUs: a sample-sharded sweep (NCCL all-reduce of [M | b | trace], collective convergence of the
refinement's conjugate gradients) must give the cores of the single-GPU sweep over all rows.  Prints one JSON line per case on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
torch.set_default_dtype(torch.float64)
import tensornetworksfork_b200 as tnb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def data(N, F, classes):
    g = torch.Generator().manual_seed(1)
    X = torch.cat([torch.rand((N, F), generator=g) * 2 - 1, torch.ones((N, 1))], 1)
    if classes:
        y = torch.nn.functional.one_hot((X[:, :F] @ torch.randn((F, classes), generator=g)).argmax(1), classes).to(torch.float64)
    else:
        y = torch.tanh(X[:, :1]) + 0.5 * X[:, 1:2] * X[:, 2:3] + 0.05 * torch.randn((N, 1), generator=g)
    return X.to(dev), y.to(dev)


def run(kind, mode, sharded, N=50001, F=6, r=8):
    classes = 3 if kind == "xe" else 0
    X, y = data(N, F, classes)
    if kind == "cpd":
        layer = tnb.CPDLayer(3, 12, F + 1, output_shape=(1,), seed=3).to(dev)
    else:
        layer = tnb.TensorTrainLayer(4, r, F + 1, output_shape=(classes - 1) if classes else 1, constrict_bond=False, seed=3).to(dev)
    tn = layer.tensor_network
    tn.gram_mode = mode
    loss = tnb.XEAutogradBregman(w=1.0) if classes else tnb.SquareBregFunction()
    if sharded:
        cut = [int(N * i / world) + (3 if 0 < i < world else 0) for i in range(world + 1)]      # uneven shards
        X, y = X[cut[rank]:cut[rank + 1]].contiguous(), y[cut[rank]:cut[rank + 1]].contiguous()
        tn.process_group, tn.shard_offset, tn.shard_total = dist.group.WORLD, cut[rank], N
    losses = []
    ok = tn.accumulating_swipe(X, y, loss, batch_size=4096, num_swipes=2, method="ridge_cholesky", eps=1.0, eps_decay=0.5,
                               loss_callback=lambda NS, n, l: losses.append(l))
    return ok, losses, [n.tensor.clone() for n in tn.train_nodes], dict(tn.solve_stats)


for kind in ("reg", "xe", "cpd"):
    for mode in ("fp64", "f16", "tf32"):
        ok_s, l_s, c_s, st = run(kind, mode, True)
        # every rank must hold the same cores bit for bit
        same = True
        for c in c_s:
            ref = c.clone()
            dist.broadcast(ref, src=0)
            same = same and bool(torch.equal(ref, c))
        flag = torch.tensor([1.0 if same else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ok_1, l_1, c_1, _ = run(kind, "fp64", False)
            core_err = max(float((a - b).norm() / b.norm()) for a, b in zip(c_s, c_1))
            loss_err = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(l_s, l_1))
            print(json.dumps({"case": kind, "gram_mode": mode, "world": world, "ok": bool(ok_s and ok_1), "ranks_bit_identical": bool(flag.item() == 1.0),
                              "core_rel_err_vs_single_gpu_fp64": core_err, "loss_rel_err": loss_err, "solve_stats": st}), flush=True)
        dist.barrier()
dist.destroy_process_group()
