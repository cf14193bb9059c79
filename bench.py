#!/usr/bin/env python
"""Benchmark of the per-site Gauss-Newton / ALS sweep (BASELINE.json metric: GN site-updates/s and samples/s per sweep).

``value`` is sample-site updates per second (rows x site updates / s) -- the whole-job aggregate that grows with the number
of GPUs under weak scaling; ``site_updates_per_s`` is reported beside it.

    python bench.py --gpus N --steps K --warmup W [--workload cfg5a] [--rows ROWS_PER_GPU] [--gram-mode tf32x3]
    python bench.py --impl reference ...        # the reference's algorithm on the host cores (oracle port)

A step is one full sweep (left-to-right + right-to-left half) of ``accumulating_swipe`` over the
rank's shard of synthetic rows (U(-1,1) features, smooth teacher target, seeds fixed).  Weak scaling:
rows per GPU are fixed, ranks shard the samples and all-reduce the per-site Gram.
Rank 0 prints ONE JSON line.  See DESIGN.md §Measurement for how every field is produced.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

if "reference" in sys.argv and os.environ.get("OMP_NUM_THREADS") == "1":
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU arm runs on rank 0 alone and is meant to use all host cores.
    # The BLAS / OpenMP pools read this when numpy / torch are imported, so it has to be lifted before the imports below.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

_REF_ROOT = None
if "reference" in sys.argv:
    # the unmodified reference (oracle/_ref, recipe oracle/make_ref.py) and its authors' environment (the opt_einsum stand-in has to
    # be importable before torch is): only the CPU arm ever imports these
    from oracle import make_ref as _make_ref
    _REF_ROOT = _make_ref.activate()

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)

# name -> model / data shape (BASELINE.json configs; SURVEY.md §8 "Config sizes")
WORKLOADS = {
    "cfg1": dict(kind="tt", sites=3, r=6, features=8, bias=True, basis=None, C=1, n=4177, constrict=True, perturb=True,
                 batch_size=512, desc="TT poly-mode abalone-shaped N=4177 F=8(+1) r=6 3 cores"),
    "cfg2": dict(kind="cpd", sites=5, r=100, features=8, bias=True, basis=None, C=1, n=20640, batch_size=-1,
                 desc="CPD rank 100 california_housing-shaped N=20640 F=8(+1) 5 factors"),
    "cfg3": dict(kind="tt", sites=90, r=24, features=90, bias=False, basis="sin-cos", C=1, n=515345, constrict=True,
                 perturb=False, batch_size=512, orthonormalize=True, desc="TNML sin-cos year-shaped N=515345 F=90 r=24 90 sites"),
    "cfg4a": dict(kind="tt", sites=784, r=38, features=784, bias=False, basis="sin-cos", C=9, n=60000, constrict=True,
                  perturb=False, batch_size=1024, solver="cg", max_iter=500, tol=1e-3,
                  desc="TNML sin-cos classifier MNIST-shaped N=60000 F=784 10 classes (C=9 logits, XE loss) r=38, matrix-free "
                       "local solve scipy_swipe('cg', max_iter=500, tol=1e-3)"),
    "cfg4b": dict(kind="conv", sites=3, r=38, CB=4, patches=50, pixels=17, features=49 * 16, bias=False, basis=None, C=9, n=60000,
                  batch_size=1024, solver="minres", max_iter=100, tol=1e-3,
                  desc="conv-TT (TensorConvolutionTrainLayer) MNIST-shaped N=60000, 49+1 patches x 16+1 pixels, 3 columns, r=38, CB=4, "
                       "9 logits (XE loss), matrix-free local solve scipy_swipe('minres', max_iter=100, tol=1e-3); P(A2)=72200"),
    "cfg5a": dict(kind="tt", sites=5, r=38, features=28, bias=True, basis=None, C=1, n=1000000, constrict=False,
                  perturb=False, batch_size=-1, desc="TT poly-mode higgs-shaped F=28(+1) r=38 5 cores (degree 5), P=41876"),
    "cfg5b": dict(kind="tt", sites=28, r=38, features=28, bias=False, basis="polynomial", degree=5, C=1, n=1000000,
                  constrict=True, perturb=False, batch_size=-1, orthonormalize=True,
                  desc="TNML polynomial degree 5 higgs-shaped F=28 r=38 28 sites, P<=8664"),
}



def site_updates_per_step(wl):
    """Site updates of one step (one sweep / one pass pair) of a workload: what both arms divide by."""
    if wl.get("solver") or wl["kind"] == "conv":
        return 2 * wl["sites"]          # matrix-free passes visit every node in both directions
    return 2 * wl["sites"] - 1          # dense sweep: the turning site is updated once


def workload_config(args, wl, rows_per_gpu, world, updates_per_step=None):
    """`config` of the JSON line: identical for the b200 arm and the reference arm run with the same arguments."""
    return {"workload": f"{args.workload}: {wl['desc']}", "rows_per_gpu": rows_per_gpu, "rows_total": rows_per_gpu * world,
            "site_updates_per_step": updates_per_step if updates_per_step is not None else site_updates_per_step(wl),
            "gram_mode": args.gram_mode, "eps": args.eps,
            "l2": "inputs larger than L2 (per-site streams of rows*(r_l+f+r_r)*8 B)", "parallelism": f"sample-shard x{world}"}

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5a", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", dest="n", type=int, default=None, help="rows per GPU (default: the workload's)")
    ap.add_argument("--gram-mode", default="f16", choices=["fp64", "tf32", "tf32x3", "f16"],
                    help="operand precision of the Gram build; in the tensor-core modes the Gram only preconditions the exact fp64 refinement "
                         "(TensorNetwork.refine = 'exact'), so 'f16' / 'tf32' (one MMA pass) give the same step as 'tf32x3' and 'fp64'")
    ap.add_argument("--flush-rows", type=int, default=None, help="fp32 accumulation window of the tensor-core Gram (default: the engine's)")
    ap.add_argument("--solve-mode", default="auto", choices=["auto", "fp64", "mixed"],
                    help="local solve: auto = tensor-core factorisation + fp64 refinement for P >= 8192 in the tf32 gram modes")
    ap.add_argument("--no-peaks", action="store_true", help="skip the cuBLAS TF32/FP64 peak measurement")
    ap.add_argument("--eps", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--max-iter", type=int, default=None, help="Krylov iterations per node (matrix-free workloads)")
    ap.add_argument("--ref-rows", type=int, default=None)
    ap.add_argument("--ref-quick", action="store_true",
                    help="reference arm as the b200 arm's cpu_baseline leg: one bounded sample (10-30 s), large solves extrapolated")
    ap.add_argument("--ref-port", action="store_true", help="reference arm on the numpy port of oracle/ even when oracle/_ref exists")
    ap.add_argument("--ref-matvecs", type=float, default=20.0,
                    help="matvecs per site assumed by --impl reference on the matrix-free workloads (the b200 arm reports its own count)")
    return ap.parse_args()


def make_data(wl, n, seed, device):
    """Synthetic rows of the workload's shape: X ~ U(-1,1), y = smooth teacher + noise."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    F = wl["features"]
    if wl["kind"] == "conv":
        # unfolded image patches with the bias patch / bias pixel of image_convolution_CG_MNIST.py:29-32
        Q, T = wl["patches"], wl["pixels"]
        X = torch.zeros((n, Q, T), dtype=torch.float64)
        X[:, :-1, :-1] = torch.rand((n, Q - 1, T - 1), generator=g, dtype=torch.float64) * 2 - 1
        X[:, -1, -1] = 1.0
        Wc = torch.randn(((Q - 1) * (T - 1), wl["C"] + 1), generator=torch.Generator(device="cpu").manual_seed(7), dtype=torch.float64)
        lab = (X[:, :-1, :-1].reshape(n, -1) @ Wc).argmax(dim=1)
        y = torch.nn.functional.one_hot(lab, num_classes=wl["C"] + 1).to(torch.float64)
        return X.to(device), y.to(device)
    X = torch.rand((n, F), generator=g, dtype=torch.float64) * 2 - 1
    w1 = torch.randn((F, 1), generator=g, dtype=torch.float64) / F ** 0.5
    w2 = torch.randn((F, 1), generator=g, dtype=torch.float64) / F ** 0.5
    y = torch.tanh(X @ w1) + 0.5 * (X @ w2) ** 2 + 0.1 * torch.randn((n, 1), generator=g, dtype=torch.float64)
    if wl["C"] > 1:     # balanced-ish classes from the argmax of a fixed random linear map, one-hot over C+1 classes
        Wc = torch.randn((F, wl["C"] + 1), generator=torch.Generator(device="cpu").manual_seed(7), dtype=torch.float64)
        y = torch.nn.functional.one_hot((X @ Wc).argmax(dim=1), num_classes=wl["C"] + 1).to(torch.float64)
    if wl["bias"]:
        X = torch.cat([X, torch.ones((n, 1), dtype=torch.float64)], dim=1)
    return X.to(device), y.to(device)


def build_model(wl, device, gram_mode):
    import tensornetworksfork_b200 as tnb
    f = wl["features"] + 1 if wl["bias"] else (2 if wl["basis"] == "sin-cos" else wl.get("degree", 3) + 1)
    if wl["kind"] == "cpd":
        layer = tnb.CPDLayer(wl["sites"], wl["r"], f, output_shape=(wl["C"],), seed=42)
    elif wl["kind"] == "conv":
        torch.manual_seed(42)
        layer = tnb.TensorConvolutionTrainLayer(wl["sites"], wl["r"], wl["patches"], wl["pixels"], wl["C"], convolution_bond=wl["CB"])
        for nd in layer.tensor_network.train_nodes:      # unit-norm random cores give vanishing logits; start at O(1) outputs
            nd.tensor = nd.tensor * 4.0
    else:
        layer = tnb.TensorTrainLayer(wl["sites"], wl["r"], f, output_shape=wl["C"], constrict_bond=wl["constrict"],
                                     perturb=wl["perturb"], seed=42)
    layer.to(device)
    layer.tensor_network.gram_mode = gram_mode
    return layer, f


def wrap_input(wl, X):
    import tensornetworksfork_b200 as tnb
    if wl["basis"] is None:
        return X
    return tnb.MappedInput(X, kind=wl["basis"], degree=wl.get("degree", 3))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class KernelTimer:
    """CUDA-event timing of individual ops.* calls on torch's current stream (the launching stream)."""

    def __init__(self, ops, names):
        self.ops = ops
        self.names = names
        self.orig = {}
        self.events = {n: [] for n in names}
        self.extra = {n: [] for n in names}
        self.enabled = False

    def install(self):
        for n in self.names:
            self.orig[n] = getattr(self.ops, n)
            setattr(self.ops, n, self._wrap(n))

    def _wrap(self, n):
        fn = self.orig[n]

        def inner(*a, **k):
            if not self.enabled:
                return fn(*a, **k)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            self.events[n].append((e0, e1))
            # keep sizes only: holding the argument tensors would pin every 14 GB system matrix of the step
            if n == "gram":
                self.extra[n].append((a[1].m, a[2].m, a[3].m, a[5]))
            elif n in ("predict", "env_update"):
                core = a[2]
                self.extra[n].append(2.0 * a[3 if n == "env_update" else 4] * core.shape[0] * core.shape[1] * core.shape[2])
            elif n == "outer_rows":      # (flops, algorithmic bytes: W streamed once, G read, out written)
                rows_, m_, ra_ = a[1].shape[0], a[1].shape[1], a[0].shape[1]
                self.extra[n].append((2.0 * rows_ * m_ * ra_, 8.0 * (rows_ * m_ + a[0].shape[0] * ra_ + ra_ * m_)))
            elif n == "rows_dot":
                rows_, m_, ra_ = a[0].shape[0], a[0].shape[1], a[1].shape[0]
                self.extra[n].append((2.0 * rows_ * m_ * ra_, 8.0 * (rows_ * m_ + rows_ * ra_ + ra_ * m_)))
            elif n == "bmm":
                A_, B_ = a[0], a[1]
                S_ = A_.shape[0] if A_.dim() == 3 else B_.shape[0]
                self.extra[n].append(2.0 * S_ * A_.shape[-2] * A_.shape[-1] * B_.shape[-1])
            elif n == "rhs":
                self.extra[n].append(2.0 * a[4] * a[0].m * a[1].m * a[2].m)
            elif n == "matvec":
                fa, fb, fc, rows = a[0], a[1], a[2], a[4]
                raw_b = 1 if fb.map_kind != 0 else fb.m
                # two passes over the three factors (8 B each), J v written and read once, weights read once
                self.extra[n].append(2 * 8.0 * rows * (fa.m / fa.div + raw_b / fb.div + fc.m / fc.div) + 3 * 8.0 * rows)
            elif n in ("cholesky_solve", "cholesky_solve_mixed", "cholesky_factor"):
                self.extra[n].append(int(a[0].shape[0]))
            elif n in ("cg", "minres", "lanczos"):
                op_ = a[0]
                by_ = 0.0
                if op_.factors is not None:
                    fa, fb, fc = op_.factors
                    raw_b = 1 if fb.map_kind != 0 else fb.m
                    # two passes over the three factors (8 B each), w*(J v) written and read once, weights read once
                    by_ = 2 * 8.0 * op_.rows * (fa.m / fa.div + raw_b / fb.div + fc.m / fc.div) + 3 * 8.0 * op_.rows
                self.extra[n].append((out[1], by_))          # the stats tensor (device): read after the timed region
            return out

        return inner

    def totals(self):
        return {n: [e0.elapsed_time(e1) for e0, e1 in ev] for n, ev in self.events.items()}


def gram_flops(call):
    """Issued multiply-adds x2 of one ops.gram call: rows * n_a*n_b*n_c pairs (the Kronecker-symmetric GEMM)."""
    ma, mb, mc, rows = call
    npair = lambda m: m * (m + 1) // 2
    P = ma * mb * mc
    return 2.0 * rows * npair(ma) * npair(mb) * npair(mc), float(rows) * P * (P + 1)


def run_sweeps(layer, x, y, wl, args, steps, counter, data_device=None):
    import tensornetworksfork_b200 as tnb
    tn = layer.tensor_network
    loss_fn = tnb.SquareBregFunction() if wl["C"] == 1 else tnb.XEAutogradBregman(w=1.0)

    def cb(NS, node):
        counter[0] += 1

    if wl.get("solver"):
        # matrix-free local solve (reference call shape: image_convolution_CG_MNIST.py:95); one step = L->R + R->L
        ok = tn.scipy_swipe(x, y, loss_fn, wl["solver"], batch_size=wl["batch_size"], num_swipes=2 * steps, lr=1.0,
                            max_iter=wl["max_iter"], tol=wl["tol"], block_callback=cb, data_device=data_device,
                            model_device=layer.tensor_network.main_nodes[0].tensor.device)
        if not ok:
            raise RuntimeError("sweep timed out")
        return
    ok = tn.accumulating_swipe(x, y, loss_fn, batch_size=wl["batch_size"], num_swipes=steps, lr=1.0, method="ridge_cholesky",
                               eps=args.eps, orthonormalize=wl.get("orthonormalize", False), block_callback=cb,
                               data_device=data_device, model_device=layer.tensor_network.main_nodes[0].tensor.device)
    if not ok:
        raise RuntimeError("sweep reported a singular system")


def bench_b200(args):
    import torch.distributed as dist
    import tensornetworksfork_b200 as tnb
    from tensornetworksfork_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL prints its version banner) must not pollute stdout: rank 0 prints exactly ONE JSON line at the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    if args.max_iter is not None and wl.get("solver"):
        wl = dict(wl, max_iter=args.max_iter)
    n = args.n if args.n is not None else wl["n"]
    layer, f = build_model(wl, dev, args.gram_mode)
    tn = layer.tensor_network
    tn.solve_mode = args.solve_mode
    if args.flush_rows is not None:
        tn.tc_flush_rows = args.flush_rows
    X, y = make_data(wl, n, seed=1000 + rank, device=dev)
    if world > 1:
        tn.process_group = dist.group.WORLD
        tn.shard_offset = rank * n
        tn.shard_total = world * n
    x = wrap_input(wl, X)
    if wl.get("orthonormalize"):
        tn.orthonormalize_left()

    timer = KernelTimer(ops, ["gram", "rhs", "env_update", "predict", "cholesky_solve", "cholesky_solve_mixed", "cholesky_factor", "cg", "minres",
                              "lanczos", "gram_trace", "gram_expand", "matvec", "outer_rows", "rows_dot", "bmm"])
    timer.install()
    counter = [0]

    # ---- ONE loop gives both numbers (the two-loop version of round 1 doubled the wall time and broke the driver's limit):
    # every step is the public call with HOST (pinned) buffers -- accumulating_swipe(x_host, y_host, model_device=cuda) copies
    # them to the device, sweeps, and the step ends with a device->host read of the fit's error.  `e2e` is the wall clock
    # over the K timed steps; `value` is the device time of the same K sweeps between the event the engine's data-ready hook
    # records once the copies are enqueued (inputs resident in HBM) and the event after the last kernel of the sweep.
    e2e_mode = not args.no_e2e
    if e2e_mode:
        Xh, yh = X.cpu().pin_memory(), y.cpu().pin_memory()
        x_head = wrap_input(wl, X[:1024].clone())
        y_head = (y[:1024] if wl["C"] == 1 else y[:1024, :wl["C"]]).clone()
        del x, X, y                       # device copies of the inputs only come from the per-step host->device copies below
        x = X = y = None
    got = []
    step_events = []
    pending = {}
    tn.on_data_ready = lambda: pending.__setitem__("e0", _rec())

    def _rec():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def one_step():
        if e2e_mode:
            run_sweeps(layer, wrap_input(wl, Xh), yh, wl, args, 1, counter, data_device=torch.device("cpu"))
        else:
            run_sweeps(layer, x, y, wl, args, 1, counter)
        step_events.append((pending.pop("e0"), _rec()))
        if e2e_mode:
            pred = tn.forward_batch(x_head, -1)
            got.append(float(((pred - y_head) ** 2).mean().item()))      # device -> host read of the result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()                       # one step = one call = one full sweep (L->R then R->L)
    barrier()
    counter[0] = 0
    step_events.clear()
    timer.enabled = True
    from tensornetworksfork_b200 import _lib as _tnlib
    launches0 = _tnlib.load().tn_launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    mv_count0 = getattr(tn, "matvec_count", 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    barrier()
    wall = time.perf_counter() - t0
    mv_count1 = getattr(tn, "matvec_count", 0)
    clk = clocks.stop() if rank == 0 else None
    timer.enabled = False
    kernel_launches = int(_tnlib.load().tn_launch_count() - launches0)
    both = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in step_events), wall * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
    ms, wall_ms = (float(v) for v in both.tolist())
    updates = counter[0]
    site_rate = updates / (ms / 1e3)               # site updates per second (does not grow with N under weak scaling)
    value = site_rate * n * world                  # sample-site updates per second: the whole-job aggregate
    e2e = None
    if e2e_mode:
        e2e = {"value": updates / (wall_ms / 1e3) * n * world, "unit": "sample-site-updates/s",
               "site_updates_per_s": updates / (wall_ms / 1e3), "ms_per_step": wall_ms / args.steps,
               "h2d_bytes_per_step": int(Xh.numel() * Xh.element_size() + yh.numel() * yh.element_size()), "d2h_bytes_per_step": 8,
               "how": "same K steps as `value`: wall clock around accumulating_swipe(x_host_pinned, y_host_pinned, model_device=cuda) "
                      "+ forward on 1024 rows + .item() of its MSE; `value` is the device time of those sweeps after the copies",
               "fit_mse_last": got[-1] if got else None}

    # ---- accuracy of the timed mode, measured (untimed) on a row subsample of this rank's data: the tensor-core Gram of the largest
    # site against the fp64 Gram kernel on the same rows, and what the refinement did inside the timed sweeps
    accuracy = None
    if args.gram_mode != "fp64" and wl["kind"] in ("tt", "cpd") and not wl.get("solver") and rank == 0:
        try:
            sub = min(n, 32768)
            xs = wrap_input(wl, (Xh[:sub] if e2e_mode else X[:sub]).to(dev))
            ys = (yh[:sub] if e2e_mode else y[:sub]).to(dev)
            tn.set_input(xs)
            tn._check_external()
            sizes = [nd.tensor.numel() for nd in tn.main_nodes]
            kbig = int(np.argmax(sizes))
            pg, tn.process_group = tn.process_group, None
            prob = tn._site_problem(kbig, ys, tnb.SquareBregFunction() if wl["C"] == 1 else tnb.XEAutogradBregman(w=1.0))
            M_tc = tn._accumulate(prob, args.gram_mode)[0].clone()
            M_64 = tn._accumulate(prob, "fp64")[0]
            tn.process_group = pg
            accuracy = {"gram_rel_err": float(((M_tc - M_64).norm() / M_64.norm()).item()), "gram_rel_err_rows": sub,
                        "gram_rel_err_site": kbig, "tc_flush_rows": tn.tc_flush_rows if tn.refine == "exact" else 2048,
                        "gram_rel_err_is": "relative Frobenius error of the unique Gram entries M, timed mode vs fp64 kernel, same rows",
                        "refine": tn.refine, "refine_rtol": tn.refine_rtol,
                        "refine_note": "refine='exact': M only preconditions conjugate gradients on the fp64 matrix-free operator; the step "
                                       "solves the reference's fp64 system to refine_max_rel (solve.stats)"}
            del M_tc, M_64, prob
        except Exception as e:          # the accuracy probe must never cost the line
            accuracy = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tot = timer.totals()
    measured = None
    if not args.no_peaks and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import peaks as _peaks
        measured = _peaks.measure()
    elif not args.no_peaks:
        # N > 1: the other ranks are waiting in destroy_process_group -- reuse the cuBLAS TF32 / FP64 peaks measured by an N = 1 run
        # on this pool (same file the N = 1 line's in-run measurement is checked against)
        try:
            measured = dict(json.load(open(os.path.join(ROOT, "profiles", "r1_peaks_tf32_fp64.json"))), source="profiles/r1_peaks_tf32_fp64.json (not re-measured at N > 1)")
        except Exception:
            measured = None
    gram_ms = tot["gram"]
    launches = kernel_launches   # kernels of libtn_b200.so launched inside the timed region (counted by the library)
    issued = algo = 0.0
    for c in timer.extra["gram"]:
        i_, a_ = gram_flops(c)
        issued += i_
        algo += a_
    gsum = sum(gram_ms) / 1e3
    if args.gram_mode == "fp64":
        if measured:
            peak, peak_src = measured["fp64_tflops_sustained"], "cuBLAS DGEMM 8192^3 measured in this run (sustained)"
        else:
            peak, peak_src = 35.5, "cuBLAS DGEMM measured on this pool earlier (profiles/r1_peaks_tf32_fp64.json)"
    elif args.gram_mode == "f16":
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = ("sustained 16-bit dense tensor peak (cuBLAS bf16 8192^3, MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long "
                    "step; fp16 and bf16 run at the same rate) -- " + ("of measured" if peaks else "of fallback 1400"))
    else:
        if measured:
            peak, peak_src = measured["tf32_tflops_sustained"], "cuBLAS TF32 GEMM 8192^3 measured in this run (sustained; kernel is timed inside a long step)"
        else:
            bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
            peak, peak_src = bf16 / 2.0, "half of the sustained bf16 dense peak of MEASURED_PEAKS.json (TF32 rate = bf16/2); " + ("of measured" if peaks else "of fallback")
    mult = 3.0 if args.gram_mode == "tf32x3" else 1.0
    issued_tf = issued * mult / gsum / 1e12 if gsum > 0 else 0.0
    # `achieved` = the multiply-adds the kernel has to EXECUTE per launch (x2), over the CUDA-event time of the Gram launches: the
    # Gram of a Kronecker Jacobian is symmetric under the swap of each factor's index pair separately, so only n_a*n_b*n_c ~ P^2/8
    # entries are distinct, 2*rows*n_a*n_b*n_c flop (x3 passes in 3xTF32) -- this is what the tensor pipe runs and what ncu's
    # sm__pipe_tensor counters see, so it is the figure set against the measured TF32 peak.  SURVEY.md 8(d) counts a plain
    # symmetric Gram, rows*P*(P+1) flop (3.67x more on this shape); that count over the same time is reported beside it as
    # `survey_equiv_tflops` -- an equivalent rate, NOT a pipe utilisation (it exceeds the TF32 peak in the one-pass mode by design).
    achieved = issued_tf
    file_peak = None
    if args.gram_mode != "fp64":
        bf16 = peaks.get("bf16_tflops_sustained")
        file_peak = (bf16 if bf16 else 1400.0) / (1.0 if args.gram_mode == "f16" else 2.0)
    traffic, traffic_note = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[f"gram_tc_kernel[{args.gram_mode}]"]
        big = max(timer.extra["gram"], key=lambda c_: gram_flops(c_)[0]) if timer.extra["gram"] else None
        if big is not None and sorted(big[:3]) == [29, 38, 38]:
            traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["rows"] * big[3]
            traffic_note = (f"dram__bytes_read+write of the dominant launch (middle site, {big[3]} rows), scaled by rows from the "
                            f"{tr['rows']}-row ncu capture ({tr['source']}); algorithmic bytes: {8 * big[3] * (38 + 29 + 38 + 1)} B of "
                            f"operands + 1.9e9 B of M written once")
    except Exception:
        pass
    roofline = {"kernel": f"gram_kr3[{args.gram_mode}]", "bound": "tensor" if args.gram_mode != "fp64" else "fp64", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_note": traffic_note,
                "achieved_is": "executed MMA flop: 2*rows*n_a*n_b*n_c unique Kronecker-pair entries" + (" x3 (hi*hi, hi*lo, lo*hi)" if mult == 3.0 else "")
                               + " / CUDA-event time of the Gram launches",
                "peak_source": peak_src,
                "frac_of_file_peak": (achieved / file_peak) if file_peak else None,
                "file_peak": file_peak, "file_peak_is": "MEASURED_PEAKS.json bf16_tflops_sustained" + ("" if args.gram_mode == "f16" else " / 2 (TF32 rate)"),
                "survey_equiv_tflops": algo / gsum / 1e12 if gsum > 0 else 0.0,
                "survey_equiv_is": "SURVEY 8(d) count rows*P*(P+1) of a plain symmetric Gram over the same time: an equivalent rate (the kernel "
                                   "skips the 3.67x of it that the Kronecker symmetry makes redundant), not a utilisation",
                "executed_flops_per_step": issued * mult / args.steps,
                "survey_flops_per_step": algo / args.steps,
                "share_of_step": gsum / (ms / 1e3), "launches": len(gram_ms), "measured_peaks": measured,
                "bf16_peaks_file": {k: peaks.get(k) for k in ("bf16_tflops", "bf16_tflops_sustained", "hbm_gbs")}}
    if wl["kind"] == "conv":
        # dominant kernels: the two passes of the matrix-free matvec over the folded patch input (streamed once each): HBM-bound
        hot = ("rows_dot", "outer_rows")
        by = sum(b_ for k_ in hot for _, b_ in timer.extra[k_])
        fl = sum(f_ for k_ in hot for f_, _ in timer.extra[k_])
        sec = sum(sum(tot[k_]) for k_ in hot) / 1e3
        hbm = peaks.get("hbm_gbs", 6550.7)
        other = ("predict", "rhs", "env_update", "bmm")
        roofline = {"kernel": "rows_dot_kernel + outer_rows_kernel (J v and J^T u passes of the matrix-free matvec over the folded patch input)",
                    "bound": "hbm", "achieved": by / sec / 1e9 if sec > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                    "frac": by / sec / 1e9 / hbm if sec > 0 else None, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6550.7 GB/s",
                    "share_of_step": sec / (ms / 1e3), "launches": sum(len(tot[k_]) for k_ in hot),
                    "fp64_tflops": fl / sec / 1e12 if sec > 0 else None,
                    "matvecs_per_site_update": (mv_count1 - mv_count0) / max(updates, 1),
                    "other_kernels_tflops": {k_: (sum(timer.extra[k_]) / (sum(tot[k_]) / 1e3) / 1e12 if tot[k_] else None) for k_ in other},
                    "measured_peaks": measured}
        wl = dict(wl, _avg_matvecs=(mv_count1 - mv_count0) / max(updates, 1))
    elif wl.get("solver"):
        # the matvecs run inside the on-device Krylov drivers (tn_cg / tn_minres): time of the driver calls, operator applications
        # from their device counters, algorithmic bytes of the two passes per application
        drv = "cg" if timer.extra["cg"] else "minres"
        st_ = [(st.tolist(), by_) for st, by_ in timer.extra[drv]]
        applies = sum(s_[3] for s_, _ in st_)
        mv_bytes = sum(s_[3] * by_ for s_, by_ in st_)
        mv_s = sum(tot[drv]) / 1e3
        hbm = peaks.get("hbm_gbs", 6550.7)
        roofline = {"kernel": f"tn_{drv}: env pass with weighted prediction epilogue + kr3 rhs pass per operator application (+ the recurrence's vector kernels)",
                    "bound": "hbm", "achieved": mv_bytes / mv_s / 1e9 if mv_s > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                    "frac": (mv_bytes / mv_s / 1e9 / hbm) if mv_s > 0 else None, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6550.7 GB/s",
                    "share_of_step": mv_s / (ms / 1e3), "launches": len(st_),
                    "matvecs_per_site_update": applies / max(updates, 1),
                    "iterations_mean": float(np.mean([s_[1] for s_, _ in st_])) if st_ else None,
                    "mean_us_per_matvec": 1e6 * mv_s / max(applies, 1), "measured_peaks": measured}
        wl = dict(wl, _avg_matvecs=applies / max(updates, 1))
    chol_ms = tot["cholesky_solve"] + tot["cholesky_solve_mixed"] + tot["cholesky_factor"]
    chol_flops = sum(P_ ** 3 / 3.0 for P_ in timer.extra["cholesky_solve"] + timer.extra["cholesky_solve_mixed"] + timer.extra["cholesky_factor"])
    n_mixed = len(tot["cholesky_solve_mixed"]) + int(tn.solve_stats.get("mixed", 0) > 0)
    cg_stats = [st.tolist() for st, _ in timer.extra["cg"]]
    refine_info = None
    if cg_stats and not wl.get("solver"):
        refine_info = {"solves": len(cg_stats), "iterations_mean": float(np.mean([c_[1] for c_ in cg_stats])),
                       "iterations_max": float(np.max([c_[1] for c_ in cg_stats])), "rel_residual_max": float(np.max([c_[0] for c_ in cg_stats])),
                       "ms_mean": float(np.mean(tot["cg"])), "share_of_step": sum(tot["cg"]) / ms,
                       "what": "tn_cg: conjugate gradients on the fp64 matrix-free operator J^T W J / sigma + ridge, preconditioned by the "
                               "Cholesky factor of the tensor-core Gram; the residual is that of the reference's fp64 system"}
    solve_info = {"kernel": "cholesky[fp64]" if n_mixed == 0 else "cholesky[3xTF32 tcgen05 trailing updates] for P >= 8192 / cholesky[fp64] for small systems",
                  "tflops_equiv": chol_flops / (sum(chol_ms) / 1e3) / 1e12 if chol_ms else None,
                  "share_of_step": (sum(chol_ms) + sum(tot["cg"])) / ms, "factor_share_of_step": sum(chol_ms) / ms,
                  "fp64_peak": measured["fp64_tflops_sustained"] if measured else 35.5,
                  "mode": tn.solve_mode, "stats": dict(tn.solve_stats), "refinement": refine_info,
                  "largest_ms": max(chol_ms) if chol_ms else None}
    shares = {k: sum(v) / ms for k, v in tot.items()}
    out = {"metric": "gn_sample_site_updates_per_s", "value": value, "unit": "sample-site-updates/s", "site_updates_per_s": site_rate,
           "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64" if args.gram_mode == "fp64" else f"f64+{args.gram_mode}", "data": "synthetic",
           "dtype_note": None if args.gram_mode == "fp64" else
           f"everything fp64 except the Gram build ({args.gram_mode} on tcgen05), which only preconditions the fp64 conjugate-gradient "
           f"refinement of every site's system: steps, cores and losses are those of the fp64 path (see `accuracy`, `solve.refinement`)",
           "config": workload_config(args, wl, n, world, updates // max(args.steps, 1)),
           "roofline": roofline, "solve": solve_info, "kernel_time_share": shares,
           "gpu_launches": launches,
           "clocks": clk, "e2e": e2e, "accuracy": accuracy}
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline is a rank-0, N = 1 figure (the driver's reference arm covers N > 1)
        out["cpu_baseline"] = cpu_baseline_leg(args, wl, n * world)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU side
def cpu_site_time(wl, rows, reps=1):
    """Seconds the oracle port needs for the Gram+rhs+forward of every distinct site shape on `rows` rows
    (one minibatch), and the dense solve per site, on all host threads."""
    from oracle import tn_oracle as orc
    import tensornetworksfork_b200 as tnb
    if wl["kind"] == "conv":
        from oracle import conv_oracle as co
        torch.manual_seed(42)
        layer = tnb.TensorConvolutionTrainLayer(wl["sites"], wl["r"], wl["patches"], wl["pixels"], wl["C"], convolution_bond=wl["CB"])
        tn_ = layer.tensor_network
        names = [nd.name for nd in tn_.train_nodes]
        cores = [nd.tensor.numpy().copy() * 4.0 for nd in tn_.train_nodes]
        rng = np.random.default_rng(0)
        X = rng.uniform(-1, 1, size=(rows, wl["patches"], wl["pixels"]))
        y = np.eye(wl["C"] + 1)[rng.integers(0, wl["C"] + 1, rows)]
        t_all = 0.0
        for idx in range(len(cores)):
            t0 = time.perf_counter()
            lo, b, parts = co.site_problem(cores, names, wl["C"], X, y, "xe", idx, -1)
            t_all += time.perf_counter() - t0
            v = rng.normal(size=b.size)
            t0 = time.perf_counter()
            co.matvec_of(parts)(v)
            t_all += (time.perf_counter() - t0) * wl.get("_avg_matvecs", 20.0)
        return t_all, 0.0, len(cores)
    f = wl["features"] + 1 if wl["bias"] else (2 if wl["basis"] == "sin-cos" else wl.get("degree", 3) + 1)
    if wl["kind"] == "cpd":
        layer = tnb.CPDLayer(wl["sites"], wl["r"], f, output_shape=(wl["C"],), seed=42)
    else:
        layer = tnb.TensorTrainLayer(wl["sites"], wl["r"], f, output_shape=wl["C"], constrict_bond=wl["constrict"], perturb=wl["perturb"], seed=42)
    cores = [n.tensor.numpy().copy() for n in layer.tensor_network.train_nodes]
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, size=(rows, wl["features"]))
    if wl["bias"]:
        X = np.concatenate([X, np.ones((rows, 1))], 1)
    y = rng.normal(size=(rows, 1))
    loss_of = orc.loss_square
    if wl["C"] > 1:
        y = np.eye(wl["C"] + 1)[rng.integers(0, wl["C"] + 1, rows)]
        loss_of = orc.loss_xe
    nmv = wl.get("_avg_matvecs")
    xin = X if wl["basis"] is None else (orc.fbasis(X) if wl["basis"] == "sin-cos" else orc.polynomial_basis(X, wl.get("degree", 3)))
    n = len(cores)
    shapes = {}
    for k in range(n):
        shapes.setdefault(cores[k].shape, []).append(k)
    t_batch = 0.0
    t_solve = 0.0
    for shp, ks in shapes.items():
        k = ks[0]
        t0 = time.perf_counter()
        if wl["kind"] == "cpd":
            pred = orc.cpd_forward(cores, xin)
            lo, g, H = orc.loss_square(pred, y)
            J = orc.cpd_jacobian(cores, xin, k)
        else:
            phis = orc.site_inputs(xin, n)
            Ls, Rs = orc.left_envs(cores, phis), orc.right_envs(cores, phis)
            pred = Ls[-1][:, :, 0]
            lo, g, H = loss_of(pred, y)
            J = orc.jacobian(cores, phis, k, Ls[k - 1] if k > 0 else np.ones((rows, 1, 1)), Rs[k + 1] if k < n - 1 else np.ones((rows, 1, 1)))
        if wl.get("solver"):
            # matrix-free local solve (network.py:770-790): rhs + nmv matvecs, each J^T (H (J v)) on the materialised batch Jacobian
            b = np.einsum("scP,sc->P", J, g)
            t_batch += (time.perf_counter() - t0) * len(ks)
            v = rng.normal(size=b.size)
            t0 = time.perf_counter()
            orc.matvec(J, H, v)
            t_batch += (time.perf_counter() - t0) * len(ks) * nmv
            continue
        A, b = orc.gram(J, g, H)
        t_batch += (time.perf_counter() - t0) * len(ks)
        P = b.size
        if P <= 4096:
            t0 = time.perf_counter()
            orc.solve_system(A, b, cores[k].ravel(), "ridge_cholesky", 1.0)
            t_solve += (time.perf_counter() - t0) * len(ks)
        else:
            t_solve += 0.0  # not timed on the bounded sample: favours the CPU number
    return t_batch, t_solve, n


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core.  Returns the thread count in effect."""
    ncpu = os.cpu_count() or 1
    try:
        # the port's arithmetic is numpy/BLAS; torch only draws the initial cores.  Its OpenMP workers spin for a while after every
        # parallel region and would steal the cores from the BLAS pool in the timed section right after: keep torch on one thread
        torch.set_num_threads(1)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits={"blas": ncpu, "openmp": 1})          # numpy's BLAS pool wide, the OpenMP pools (torch) out of its way
    except Exception:
        pass
    return ncpu


def wake_host_cores(max_seconds=8.0):
    """Multi-threaded GEMMs until their time settles.  After an idle stretch (the host idles while the GPU arm runs) the first
    seconds of multi-threaded BLAS on these virtualised hosts run 20-400x slow (measured: every call of a fresh process ~0.8 s
    instead of ~2 ms, until the vCPUs are back); a CPU number taken then says nothing about the CPU."""
    a = np.random.default_rng(0).normal(size=(1024, 1024))
    t_start = time.perf_counter()
    prev = None
    while time.perf_counter() - t_start < max_seconds:
        t0 = time.perf_counter()
        for _ in range(4):
            a @ a
        dt = time.perf_counter() - t0
        if prev is not None and dt < 0.25 and abs(dt - prev) < 0.2 * prev and time.perf_counter() - t_start > 1.0:
            break
        prev = dt


def cpu_site_time_warm(wl, rows):
    """cpu_site_time without first-call effects (BLAS / LAPACK thread pools, lazy imports: ~1 s, 50x the cost of a small
    workload): a short untimed call first, and the faster of two timed calls when one call is cheap."""
    wake_host_cores()
    cpu_site_time(wl, min(rows, 64))
    t0 = time.perf_counter()
    best = cpu_site_time(wl, rows)
    if time.perf_counter() - t0 < 10.0:
        again = cpu_site_time(wl, rows)
        if again[0] + again[1] < best[0] + best[1]:
            best = again
    return best


def cpu_baseline_leg(args, wl, rows_total):
    """cpu_baseline of the b200 line: the unmodified reference (oracle/_ref) on a bounded sample in a child process (its opt_einsum
    stand-in has to be on sys.path before torch is imported, and its thread pools should not inherit this process's), else the port."""
    from oracle import make_ref
    if make_ref.ref_root() is not None and not wl.get("solver") and wl["kind"] in ("tt", "cpd"):
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS", "MKL_NUM_THREADS")}
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", args.workload, "--steps", "1",
               "--warmup", "1", "--ref-quick", "--rows", str(rows_total), "--eps", str(args.eps)]
        if args.ref_rows:
            cmd += ["--ref-rows", str(args.ref_rows)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)   # bounded: the whole N = 1 line must stay far inside the driver's per-run limit
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            return json.loads(line)["cpu_baseline"]
        except Exception as e:      # fall through to the port, and say why
            base = cpu_baseline(args, wl, rows_total)
            base["sample"] += f" [reference child failed: {type(e).__name__}]"
            return base
    return cpu_baseline(args, wl, rows_total)


def cpu_baseline(args, wl, rows_total):
    use_all_host_threads()
    rows = args.ref_rows or (256 if args.workload == "cfg5a" else ((32 if wl["kind"] == "conv" else 128) if wl.get("solver") else 2048))
    if wl.get("solver") and "_avg_matvecs" not in wl:
        wl = dict(wl, _avg_matvecs=float(args.ref_matvecs))
    t_batch, t_solve, n = cpu_site_time_warm(wl, rows)
    per_sweep_sites = max(2 * n - 2, 1)
    t_all_sites = t_batch * (rows_total / rows) + t_solve     # every site once
    value = n / t_all_sites * rows_total
    if wl.get("solver"):
        return {"value": value, "unit": "sample-site-updates/s", "site_updates_per_s": n / t_all_sites, "cores": os.cpu_count(),
                "kind": "port",
                "sample": f"oracle port (numpy/BLAS, all host threads) of the matrix-free site update: envs + batch Jacobian + rhs + "
                          f"{wl['_avg_matvecs']:.1f} matvecs J^T(H(Jv)) per site (the count the GPU run needed on average) for every "
                          f"distinct site shape on one {rows}-row minibatch ({t_batch:.2f} s for all sites); per-row cost "
                          f"extrapolated linearly to {rows_total} rows"}
    return {"value": value, "unit": "sample-site-updates/s", "site_updates_per_s": n / t_all_sites, "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle port (numpy/BLAS, all host threads): env + Jacobian + Gram + rhs of every site on one {rows}-row "
                      f"minibatch ({t_batch:.2f} s) plus the dense solves with P<=4096 ({t_solve:.2f} s; larger P not timed, "
                      f"which favours the CPU); per-row cost extrapolated linearly to {rows_total} rows.  Calibration of the port "
                      f"against the unmodified reference on one 8-core host (profiles/r1_reference_vs_port_cpu.json): 23-48x faster "
                      f"than the reference verbatim (S x P x P einsum temporary), 1.1-2.8x slower than the reference with opt_einsum"}


# ------------------------------------------------------------------------------------------ CPU side: the unmodified reference
def _ref_model(wl, rows):
    """The reference's own layer + inputs for a workload (tensor/layers.py of oracle/_ref), on `rows` synthetic rows."""
    from tensor.layers import TensorTrainLayer, CPDLayer          # the REFERENCE's modules (oracle/_ref on sys.path)
    from tensor.bregman import SquareBregFunction
    f = wl["features"] + 1 if wl["bias"] else (2 if wl["basis"] == "sin-cos" else wl.get("degree", 3) + 1)
    if wl["kind"] == "cpd":
        layer = CPDLayer(wl["sites"], wl["r"], f, output_shape=(wl["C"],), seed=42)
    else:
        layer = TensorTrainLayer(wl["sites"], wl["r"], f, output_shape=wl["C"], constrict_bond=wl["constrict"], perturb=wl["perturb"], seed=42)
    X, y = make_data(wl, rows, seed=1000, device="cpu")
    if wl["basis"] == "sin-cos":
        x = [torch.stack([torch.cos(0.5 * np.pi * X[:, j]), torch.sin(0.5 * np.pi * X[:, j])], 1) for j in range(X.shape[1])]
    elif wl["basis"] == "polynomial":
        x = [torch.stack([X[:, j] ** d for d in range(wl.get("degree", 3) + 1)], 1) for j in range(X.shape[1])]
    else:
        x = X
    tn = layer.tensor_network
    if wl.get("orthonormalize"):
        tn.orthonormalize_left()
    return tn, x, y, SquareBregFunction()


class _SolveTimer:
    """Instrumentation around the reference's own `solve_system` (instance attribute; the class is untouched): times every call and,
    for systems above `skip_above`, returns a zero step instead of solving (their cost is measured once, separately)."""

    def __init__(self, tn, skip_above):
        self.orig = tn.solve_system
        self.skip_above = skip_above
        self.seconds = 0.0
        self.skipped = []
        tn.solve_system = self

    def __call__(self, node, A, b, method="exact", eps=0.0):
        P = b.numel()
        if P > self.skip_above:
            self.skipped.append(P)
            return torch.zeros_like(b)
        t0 = time.perf_counter()
        out = self.orig(node, A, b, method=method, eps=eps)
        self.seconds += time.perf_counter() - t0
        return out


def _ref_solve_seconds(tn, P, eps, budget_s, note):
    """Seconds of the reference's own solve_system(ridge_cholesky) at size P on the host cores: measured when the memory and the
    time budget allow (predicted from a P = 6144 solve, cubic), else that cubic extrapolation -- `note` says which."""
    node = tn.train_nodes[0]

    def run(Pq):
        g = torch.Generator().manual_seed(Pq)
        Bm = torch.randn((Pq, 64), generator=g)
        A = Bm @ Bm.t()
        A.diagonal().add_(float(Pq))
        b = torch.randn((Pq,), generator=g)

        class _N:                      # solve_system only reads node.tensor for the ridge term
            tensor = torch.zeros((Pq,))
        t0 = time.perf_counter()
        tn.solve_system(_N, A, b, method="ridge_cholesky", eps=eps)
        return time.perf_counter() - t0

    Pq = min(P, 6144)
    run(min(Pq, 1024))
    tq = run(Pq)
    if Pq == P:
        note.append(f"solve_system at P={P} measured: {tq:.2f} s")
        return tq
    pred = tq * (P / Pq) ** 3
    avail = None
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        pass
    need = 5.5 * 8.0 * P * P               # A, A/scale, eye, A+ridge, L (network.py:296-315)
    if pred <= budget_s and avail is not None and avail > 1.3 * need:
        t = run(P)
        note.append(f"solve_system at P={P} measured once: {t:.1f} s (cubic prediction from P={Pq}: {pred:.1f} s)")
        return t
    note.append(f"solve_system at P={P} NOT run ({'predicted %.0f s > budget %.0f s' % (pred, budget_s) if pred > budget_s else 'host memory'}): "
                f"cubic extrapolation {pred:.1f} s from the measured {tq:.2f} s at P={Pq}")
    return pred


def reference_arm(args, wl, n_total, steps, warmup, quick):
    """Sample-site updates/s of the UNMODIFIED reference (oracle/_ref: tensor/network.py:379-608) on the host cores.

    A step = the reference's own accumulating_swipe(node_order=[node], skip_second=True, ...) -- forward, loss, get_A_b, solve_system,
    update_node -- for every distinct core shape of the train, on one `rows`-row sample of the workload's data.  Time per sweep at
    N rows = sum over the sweep's site updates of [t_sample(shape) * N / rows + t_solve(shape)]: everything but the solve is linear in
    the rows (the reference loops over minibatches), the solve is not; solves above P = 4096 are stubbed out of the sample (zero
    step) and their cost is measured once by calling the reference's solve_system at that size."""
    from tensornetworksfork_b200.tensor.network import sweep_schedule
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    rows = args.ref_rows or (256 if args.workload == "cfg5a" else 2048)
    tn, x, y, loss_fn = _ref_model(wl, rows)
    nodes = list(tn.train_nodes)
    shapes = {}
    for k, nd in enumerate(nodes):
        shapes.setdefault(tuple(nd.tensor.shape), k)
    sched = sweep_schedule(list(range(len(nodes))), list(reversed(range(len(nodes)))), 1, args.eps)
    visits = [nodes[(i if half == 0 else len(nodes) - 1 - i)] for _, half, i, _ in sched]
    count = {shp: sum(1 for nd in visits if tuple(nd.tensor.shape) == shp) for shp in shapes}
    st = _SolveTimer(tn, skip_above=4096)
    note = []

    def sample():
        per = {}
        for shp, k in shapes.items():
            s0 = st.seconds
            t0 = time.perf_counter()
            ok = tn.accumulating_swipe(x, y, loss_fn, node_order=[nodes[k]], batch_size=-1, num_swipes=1, skip_second=True,
                                       method="ridge_cholesky", eps=args.eps, orthonormalize=False)
            dt = time.perf_counter() - t0
            assert ok
            per[shp] = (dt - (st.seconds - s0), st.seconds - s0)        # (linear-in-rows part, small solve)
        return per

    tn.solve_system = st.orig
    big = sorted({int(np.prod(shp)) for shp in shapes if int(np.prod(shp)) > 4096})
    t_big = {P: _ref_solve_seconds(tn, P, args.eps, 0.0 if quick else 150.0, note) for P in big}
    tn.solve_system = st
    wake_host_cores()
    for _ in range(warmup):
        sample()
    vals, walls = [], []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        per = sample()
        walls.append(time.perf_counter() - t0)
        t_sweep = 0.0
        for shp, (t_lin, t_small) in per.items():
            P = int(np.prod(shp))
            t_sweep += count[shp] * (t_lin * (n_total / rows) + (t_big[P] if P in t_big else t_small))
        vals.append(len(visits) * n_total / t_sweep)
    value = float(np.median(vals))
    sample_txt = (f"UNMODIFIED reference (oracle/_ref = tensor/ + models/ of the reference tree, opt_einsum stand-in as in its authors' "
                  f"environment; torch {torch.__version__}, {torch.get_num_threads()} threads of {ncpu} host cores): per step its own "
                  f"accumulating_swipe(node_order=[node], skip_second=True, method='ridge_cholesky') on one {rows}-row sample for each of the "
                  f"{len(shapes)} distinct core shapes ({float(np.median(walls)):.2f} s per step); a sweep of {len(visits)} site updates at "
                  f"{n_total} rows = sample time x rows/{rows} + the solves; " + "; ".join(note))
    return value, {"value": value, "unit": "sample-site-updates/s", "site_updates_per_s": value / n_total, "cores": ncpu,
                   "kind": "reference", "sample": sample_txt, "opt_einsum_path": bool(torch.backends.opt_einsum.is_available())}, float(np.mean(walls))


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = (args.n if args.n is not None else wl["n"]) * world
    use_ref = _REF_ROOT is not None and not args.ref_port and not wl.get("solver") and wl["kind"] in ("tt", "cpd")
    if use_ref:
        value, base, step_s = reference_arm(args, wl, n, args.steps, args.warmup, args.ref_quick)
    else:
        value, base, step_s = port_arm(args, wl, n)
    out = {"impl": "reference", "metric": "gn_sample_site_updates_per_s", "value": value, "unit": "sample-site-updates/s",
           "site_updates_per_s": value / n, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, wl, n // world, world),      # the b200 arm's config, key for key (the host side is in cpu_baseline)
           "cpu_baseline": base,
           "e2e": {"value": value, "unit": "sample-site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def port_arm(args, wl, n):
    """The numpy port of oracle/ on the host cores: used where oracle/_ref is absent and for the matrix-free workloads."""
    use_all_host_threads()
    rows = args.ref_rows or (256 if args.workload == "cfg5a" else ((32 if wl["kind"] == "conv" else 128) if wl.get("solver") else 2048))
    if wl.get("solver"):
        wl = dict(wl, _avg_matvecs=float(args.ref_matvecs))
    for _ in range(args.warmup):
        cpu_site_time(wl, min(rows, 64))
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(args.steps, 1)):
        t_batch, t_solve, ns = cpu_site_time_warm(wl, rows) if not vals else cpu_site_time(wl, rows)
        vals.append(ns / (t_batch * (n / rows) + t_solve) * n)
    wall = time.perf_counter() - t0
    value = float(np.median(vals))
    sample = (f"oracle port of tensor/network.py (numpy/BLAS, {os.cpu_count()} host threads): per step, env+Jacobian+Gram+rhs of every "
              f"site on one {rows}-row minibatch and the dense solves with P<=4096, extrapolated linearly to {n} rows (oracle/_ref not "
              f"present or workload not covered by the reference arm)")
    if wl.get("solver"):
        sample = (f"oracle port of tensor/network.py:709-932 (numpy/BLAS, {os.cpu_count()} host threads): per step, envs + batch Jacobian + "
                  f"rhs + {wl['_avg_matvecs']:.0f} matvecs per site (--ref-matvecs) of every distinct site shape on one {rows}-row "
                  f"minibatch, extrapolated linearly to {n} rows")
    return value, {"value": value, "unit": "sample-site-updates/s", "cores": os.cpu_count(), "kind": "port", "sample": sample}, wall / max(args.steps, 1)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        bench_reference(a)
    else:
        bench_b200(a)
