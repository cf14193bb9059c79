"""Pins oracle/tn_oracle.py against recordings of the reference itself (CPU only)."""
import numpy as np
import pytest

import golden_util as gu
from oracle import tn_oracle as orc


def _site(fx, u, cores):
    meta = fx["meta"]
    kw = dict(loss=meta["loss"], batch_size=meta["batch_size"], method=u.get("method", meta["method"]),
              eps=u["eps"], lr=meta["lr"], apply=False, adaptive_step=meta.get("adaptive_step", False), max_norm=meta.get("max_norm"))
    if meta["loss"] == "xe":
        kw["loss_kwargs"] = {"w": meta["w"]}
    m = meta["method"]
    if m == "ridge_exact" and u["eps"] == 0:
        m = "exact"
    kw["method"] = m
    if meta["kind"] == "cpd":
        kw.pop("adaptive_step"), kw.pop("max_norm")
        return orc.cpd_site_update(cores, fx["x"], fx["y"], u["k"], **kw)
    return orc.site_update(cores, fx["x"], fx["y"], u["k"], **kw)


@pytest.mark.parametrize("name", gu.names())
def test_teacher_forced_site_updates(name):
    """Same cores, same data -> A, b, loss, step, new core as the reference."""
    fx = gu.load(name)
    meta = fx["meta"]
    for u in fx["updates"]:
        cores = [c.copy() for c in u["before"]]
        r = _site(fx, u, cores)
        assert gu.relerr(r["A"].reshape(u["A"].reshape(r["A"].shape).shape), u["A"].reshape(r["A"].shape)) < 1e-12
        assert gu.relerr(r["b"].ravel(), u["b"].ravel()) < 1e-12
        assert abs(r["loss"] - u["loss"]) <= 1e-12 * max(1.0, abs(u["loss"]))
        # step: compare through the backward error of the reference's own linear system
        P = r["b"].size
        A = u["A"].reshape(P, P)
        sc = np.abs(np.diag(A)).mean() or 1.0
        ridge = 0.0 if meta["method"] in ("exact", "cholesky") else 2 * u["eps"]
        M = A / sc + ridge * np.eye(P)
        rhs = u["b"].ravel() / sc + ridge * u["before"][u["k"]].ravel()
        res = np.linalg.norm(M @ r["step"].ravel() + rhs) / max(np.linalg.norm(rhs), 1e-300)
        res_ref = np.linalg.norm(M @ u["step"].ravel() + rhs) / max(np.linalg.norm(rhs), 1e-300)
        assert res <= max(10 * res_ref, 1e-10)
        cond = np.linalg.cond(M)
        assert gu.relerr(r["step"].ravel(), u["step"].ravel()) < 1e-13 * cond + 1e-12
        if not meta.get("orthonormalize"):
            new = orc.update_node(u["before"][u["k"]], u["step"], lr=meta["lr"], adaptive_step=meta.get("adaptive_step", False),
                                  max_norm=meta.get("max_norm"))
            assert gu.relerr(new, u["after"][u["k"]]) < 1e-13


@pytest.mark.parametrize("name", [n for n in gu.names() if "cpd" not in n])
def test_environments(name):
    fx = gu.load(name)
    n = len(fx["cores0"])
    meta = fx["meta"]
    bs = meta["batch_size"]
    for u in fx["updates"]:
        if "L" not in u and "R" not in u:
            continue
        phis = orc.site_inputs(fx["x"], n)
        N = phis[0].shape[0]
        b = N if bs <= 0 else bs
        sl = slice(0, min(b, N))  # the recording keeps the stacks of the FIRST minibatch
        ph = [p[sl] for p in phis]
        Ls = orc.left_envs(u["before"], ph)
        Rs = orc.right_envs(u["before"], ph)
        k = u["k"]
        if "L" in u and k > 0:
            assert gu.relerr(Ls[k - 1], u["L"].reshape(Ls[k - 1].shape)) < 1e-13
        if "R" in u and k < n - 1:
            assert gu.relerr(Rs[k + 1], u["R"].reshape(Rs[k + 1].shape)) < 1e-13


@pytest.mark.parametrize("name", [n for n in gu.names() if "cpd" not in n])
def test_free_running_sweep(name):
    """Site order, eps schedule, turn-around skip, QR re-gauge: the whole driver.
    Settings in the fixtures are well conditioned (eps >= 0.1), SURVEY.md §7.3 item 3."""
    fx = gu.load(name)
    meta = fx["meta"]
    cores = [c.copy() for c in fx["cores0"]]
    trace = []
    kw = {}
    if meta["loss"] == "xe":
        kw["loss_kwargs"] = {"w": meta["w"]}
    ok = orc.accumulating_swipe(cores, fx["x"], fx["y"], loss=meta["loss"], batch_size=meta["batch_size"],
                                num_swipes=meta["num_swipes"], lr=meta["lr"], method=meta["method"], eps=meta["eps"],
                                eps_decay=meta.get("eps_decay"), orthonormalize=meta.get("orthonormalize", False),
                                skip_second=meta.get("skip_second", False), trace=trace, **gu.sweep_extras(meta), **kw)
    assert ok == fx["ok"]
    assert [(t["NS"], t["k"]) for t in trace] == [(u["NS"], u["k"]) for u in fx["updates"]]
    for t, u in zip(trace, fx["updates"]):
        assert abs(t["eps"] - u["eps"]) <= 1e-15 * max(1, abs(u["eps"]))
        assert abs(t["loss"] - u["loss"]) <= 1e-8 * max(1.0, abs(u["loss"]))
    if meta["method"] == "exact":
        return  # unregularised LU of a gauge-singular system: trajectories are not comparable
    pred = orc.forward(cores, fx["x"])
    assert gu.relerr(pred, fx["pred"].reshape(pred.shape)) < 1e-8
    if meta.get("orthonormalize"):
        for c, ref in zip(cores, fx["updates"][-1]["after"]):
            assert gu.relerr(c, ref) < 1e-7


def test_cpd_forward():
    fx = gu.load("cpd_reg")
    last = fx["updates"][-1]["after"]
    pred = orc.cpd_forward(last, fx["x"])
    assert gu.relerr(pred, fx["pred"].reshape(pred.shape)) < 1e-12


def test_lanczos_and_matvec_consistency():
    rng = np.random.default_rng(0)
    J = rng.normal(size=(50, 2, 7))
    Hh = rng.normal(size=(50, 2, 2))
    H = np.einsum("sij,skj->sik", Hh, Hh)
    A, _ = orc.gram(J, np.zeros((50, 2)), H)
    v = rng.normal(size=7)
    assert gu.relerr(orc.matvec(J, H, v), A @ v) < 1e-12
    b = rng.normal(size=7)
    x = orc.lanczos_solve(lambda t: A @ t, b, np.zeros(7), 7, 1e-14)
    assert gu.relerr(A @ x, b) < 1e-8
