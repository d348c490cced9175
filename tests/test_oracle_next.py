"""Pins of oracle/nodes_next.py (SURVEY.md 8-f nodes), CPU only.

Three independent pins:
  1. the literal vectors of the reference's own node tests (tests/golden/tensors_next.json, lifted by
     tests/golden/make_goldens.py) at the reference's tolerance;
  2. the closed forms the reference's ENABLED tests state as code (exp/test.rs:24-36, logn/test.rs:24-36,
     multiplication/test.rs, division/test.rs, subtraction/test.rs, pad/{reflective,replicative}/test.rs);
  3. torch-CPU autograd / torch.optim on random, non-symmetric data.
"""
import json
import os

import numpy as np
import pytest

import oracle as O

F32 = np.float32
EPS = 4.88e-4          # the reference's assert_almost_equals / are_similar tolerance
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    with open(os.path.join(HERE, "golden", "tensors_next.json")) as fh:
        return json.load(fh)


def T(e):
    return np.asarray(e["values"], F32).reshape(e["shape"])


UNARY = [("negation", "neg", 0), ("sqrt", "sqrt", 0), ("sigmoid", "sigmoid", 0), ("tanh", "tanh", 0),
         ("softplus", "softplus", 0), ("leaky_relu", "leaky_relu", 0), ("power", "powi", 3)]


@pytest.mark.parametrize("file,op,ip", UNARY)
def test_unary_forward_goldens(G, file, op, ip):
    """forward(): tensors = [input, expected, input+1, expected (no re-evaluation), expected of input+1]"""
    t = G[file]["forward"][0]["tensors"]
    assert np.allclose(O.unary_forward(op, T(t[0]), ip), T(t[1]), atol=EPS, rtol=1e-4)
    assert np.allclose(O.unary_forward(op, T(t[2]), ip), T(t[4]), atol=EPS, rtol=1e-4)


@pytest.mark.parametrize("file,op,ip", [u for u in UNARY if u[0] != "negation"])
def test_unary_backward_goldens(G, file, op, ip):
    """backward(): tensors = [zeros, input x, seed, seed, after 1st backward, after 2nd (accumulated), overwrite]"""
    t = G[file]["backward"][0]["tensors"]
    x, g = T(t[1]), T(t[2])
    # sigmoid / tanh build the forward node from x; sqrt/test.rs:170 hands the node its OUTPUT (1, 1.4142, 1.7321)
    saved = O.unary_forward(op, x, ip) if op in ("sigmoid", "tanh") else x
    dx = np.zeros_like(x)
    O.unary_backward(op, g, saved, dx, ip)
    assert np.allclose(dx, T(t[4]), atol=EPS, rtol=1e-4)
    O.unary_backward(op, g, saved, dx, ip)
    assert np.allclose(dx, T(t[5]), atol=2 * EPS, rtol=1e-4)


def test_power_negative_exponent_golden(G):
    t = G["power"]["backward_negative_exp"][0]["tensors"]          # power/test.rs:195-225, exponent -3
    dx = np.zeros(3, F32)
    O.unary_backward("powi", T(t[2]), T(t[1]), dx, -3)
    assert np.allclose(dx, T(t[4]), atol=EPS)


def test_negation_and_transpose_goldens(G):
    t = G["negation"]["backward"][0]["tensors"]                     # [zeros, seed, seed, -1s, -2s, -1s]
    dx = np.zeros(T(t[0]).shape, F32)
    O.unary_backward("neg", T(t[1]), None, dx)
    assert np.allclose(dx, T(t[3]), atol=EPS)
    O.unary_backward("neg", T(t[1]), None, dx)
    assert np.allclose(dx, T(t[4]), atol=EPS)
    f = G["transpose"]["forward"][0]["tensors"]
    assert np.array_equal(O.transpose_forward(T(f[0])), T(f[1]))
    b = G["transpose"]["backward"][0]["tensors"]                    # diff (4,3), seed (3,4)
    dx = np.zeros((4, 3), F32)
    O.transpose_backward(T(b[1]), dx)
    assert np.array_equal(dx, T(b[3]))
    O.transpose_backward(T(b[1]), dx)
    assert np.array_equal(dx, T(b[4]))


def test_mv_vm_vv_goldens(G):
    f = G["matrix_vector_mul"]["forward"][0]["tensors"]             # [A, v, y, ...]
    assert np.allclose(O.mv_forward(T(f[0]), T(f[1])), T(f[2]), atol=EPS)
    b = G["matrix_vector_mul"]["backward"][0]["tensors"]            # [dA0, dv0, A, v, seed, seed, dA, dv, 2dA, 2dv, ...]
    a, v, g = T(b[2]), T(b[3]), T(b[4])
    da, dv = np.zeros_like(a), np.zeros_like(v)
    O.mv_backward(a, v, g, da, dv)
    assert np.allclose(da, T(b[6]), atol=EPS) and np.allclose(dv, T(b[7]), atol=EPS)
    O.mv_backward(a, v, g, da, dv)
    assert np.allclose(da, T(b[8]), atol=EPS) and np.allclose(dv, T(b[9]), atol=EPS)
    f = G["vector_matrix_mul"]["forward"][0]["tensors"]             # [v, A, y, ...]
    assert np.allclose(O.vm_forward(T(f[0]), T(f[1])), T(f[2]), atol=EPS)
    b = G["vector_matrix_mul"]["backward"][0]["tensors"]            # [dv0, dA0, v, A, seed, seed, dv, dA, ...]
    v, a, g = T(b[2]), T(b[3]), T(b[4])
    dv, da = np.zeros_like(v), np.zeros_like(a)
    O.vm_backward(v, a, g, dv, da)
    assert np.allclose(dv, T(b[6]), atol=EPS) and np.allclose(da, T(b[7]), atol=EPS)
    f = G["vector_vector_mul"]["forward"][0]
    assert abs(float(O.vv_forward(T(f["tensors"][0]), T(f["tensors"][1]))) - f["scalars"][0]) < EPS   # arr0(12.0)
    b = G["vector_vector_mul"]["backward"][0]["tensors"]            # [dl0, dr0, l, r, dl, dr, 2dl, 2dr, dl, dr]
    l, r = T(b[2]), T(b[3])
    dl, dr = np.zeros_like(l), np.zeros_like(r)
    O.vv_backward(l, r, np.float32(1.0), dl, dr)
    assert np.allclose(dl, T(b[4]), atol=EPS) and np.allclose(dr, T(b[5]), atol=EPS)


def test_closed_forms_of_the_enabled_reference_tests():
    x = np.linspace(-4, 4, 9, dtype=F32).reshape(3, 3)
    want = np.array([np.exp(F32(-4 + i)) for i in range(9)], F32).reshape(3, 3)      # exp/test.rs:24-36
    assert np.allclose(O.unary_forward("exp", x), want, rtol=1e-6)
    dx = np.zeros((3, 3), F32)
    O.unary_backward("exp", np.ones((3, 3), F32), want, dx)                          # exp/test.rs:59-73
    O.unary_backward("exp", np.ones((3, 3), F32), want, dx)
    assert np.allclose(dx, want * 2, rtol=1e-6)
    xp = np.linspace(1, 9, 9, dtype=F32).reshape(3, 3)
    assert np.allclose(O.unary_forward("ln", xp), np.log(xp), rtol=1e-6)             # logn/test.rs
    # multiplication / division / subtraction with a (3,) operand broadcast against (3,3)
    l, r = np.linspace(1, 9, 9, dtype=F32).reshape(3, 3), np.array([1, 2, 3], F32)
    g = np.ones((3, 3), F32)
    for op, fwd in (("mul", l * r), ("div", l / r), ("sub", l - r)):
        assert np.allclose(O.binary_forward(op, l, r), fwd, rtol=1e-6)
    dl, dr = np.zeros_like(l), np.zeros_like(r)
    O.binary_backward("mul", g, l, r, dl, dr)
    assert np.allclose(dl, np.broadcast_to(r, (3, 3))) and np.allclose(dr, l.sum(0))
    dl, dr = np.zeros_like(l), np.zeros_like(r)
    O.binary_backward("div", g, l, r, dl, dr)
    assert np.allclose(dl, 1 / np.broadcast_to(r, (3, 3))) and np.allclose(dr, (-l / r ** 2).sum(0), rtol=1e-6)
    dl, dr = np.zeros_like(l), np.zeros_like(r)
    O.binary_backward("sub", g, l, r, dl, dr)
    assert np.allclose(dl, 1) and np.allclose(dr, -3)


def test_pad_modes_match_the_reference_index_maps():
    """pad/reflective/mod.rs:37-80 and pad/replicative/mod.rs:37-80 restated literally (double loop) vs the oracle"""
    rng = np.random.default_rng(0)
    base = rng.standard_normal((4, 5)).astype(F32)
    for mode in ("reflective", "replicative"):
        for px, py in ((1, 2), (2, 0), (3, 4) if mode == "replicative" else (3, 3)):
            lx, ly = base.shape
            out = np.zeros((lx + 2 * px, ly + 2 * py), F32)
            for i in range(lx + 2 * px):
                for j in range(ly + 2 * py):
                    if mode == "reflective":
                        qx = (py * 2 - j if j < py else (j if j < ly + py else (ly + py - 1) * 2 - j)) - py
                        qy = (px * 2 - i if i < px else (i if i < lx + px else (lx + px - 1) * 2 - i)) - px
                    else:
                        qx = (py if j < py else (j if j < ly + py else ly + py - 1)) - py
                        qy = (px if i < px else (i if i < lx + px else lx + px - 1)) - px
                    out[i, j] = base[qy, qx]
            got = O.pad_mode_forward(base[None, None], (px, py), mode)[0, 0]
            assert np.array_equal(got, out), (mode, px, py)
    # numpy's own modes agree (reflect / edge), all three ranks
    for shape, pad in (((2, 3, 7), (2,)), ((2, 2, 4, 5), (1, 3)), ((1, 2, 3, 4, 5), (2, 1, 3))):
        x = rng.standard_normal(shape).astype(F32)
        width = [(0, 0), (0, 0)] + [(p, p) for p in pad]
        assert np.array_equal(O.pad_mode_forward(x, pad, "reflective"), np.pad(x, width, mode="reflect"))
        assert np.array_equal(O.pad_mode_forward(x, pad, "replicative"), np.pad(x, width, mode="edge"))
        assert np.array_equal(O.pad_mode_forward(x, pad, "constant", 1.5), np.pad(x, width, constant_values=1.5))
        g = rng.standard_normal(O.pad_mode_forward(x, pad, "constant").shape).astype(F32)
        dx = np.ones(shape, F32)
        O.pad_mode_backward(g, dx, pad)
        sl = tuple([slice(None), slice(None)] + [slice(p, -p if p else None) for p in pad])
        assert np.array_equal(dx, 1 + g[sl])


# ------------------------------------------------------------------------------------------- torch cross-checks
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("op", ["sub", "mul", "div"])
@pytest.mark.parametrize("ls,rs", [((5, 7), (5, 7)), ((4, 1, 6), (3, 6)), ((6,), (2, 3, 6)), ((3, 1), (1, 4))])
def test_binary_against_torch(op, ls, rs):
    rng = np.random.default_rng(hash((op, ls, rs)) % 2 ** 31)
    l = rng.uniform(0.5, 2, ls).astype(F32)
    r = rng.uniform(0.5, 2, rs).astype(F32)
    tl, tr = torch.tensor(l, requires_grad=True), torch.tensor(r, requires_grad=True)
    ty = {"sub": tl - tr, "mul": tl * tr, "div": tl / tr}[op]
    g = rng.standard_normal(tuple(ty.shape)).astype(F32)
    ty.backward(torch.tensor(g))
    assert np.allclose(O.binary_forward(op, l, r), ty.detach().numpy(), rtol=1e-6, atol=1e-6)
    dl, dr = np.zeros_like(l), np.zeros_like(r)
    O.binary_backward(op, g, l, r, dl, dr)
    assert np.allclose(dl, tl.grad.numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose(dr, tr.grad.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("op,ip", [("neg", 0), ("exp", 0), ("ln", 0), ("sqrt", 0), ("sigmoid", 0), ("tanh", 0),
                                   ("softplus", 0), ("leaky_relu", 0), ("powi", 3), ("powi", -2), ("powi", 0)])
def test_unary_against_torch(op, ip):
    rng = np.random.default_rng(7)
    lo = 0.2 if op in ("ln", "sqrt") or (op == "powi" and ip < 0) else -2.0
    x = rng.uniform(lo, 2.0, (6, 9)).astype(F32)
    tx = torch.tensor(x, requires_grad=True)
    fn = {"neg": lambda t: -t, "exp": torch.exp, "ln": torch.log, "sqrt": torch.sqrt, "sigmoid": torch.sigmoid,
          "tanh": torch.tanh, "softplus": torch.nn.functional.softplus,
          "leaky_relu": lambda t: torch.nn.functional.leaky_relu(t, 0.01), "powi": lambda t: t ** ip}[op]
    ty = fn(tx)
    g = rng.standard_normal(x.shape).astype(F32)
    ty.backward(torch.tensor(g))
    y = O.unary_forward(op, x, ip)
    assert np.allclose(y, ty.detach().numpy(), rtol=1e-5, atol=1e-6)
    dx = np.zeros_like(x)
    O.unary_backward(op, g, y if op in O.UNARY_SAVES_OUTPUT else x, dx, ip)
    assert np.allclose(dx, tx.grad.numpy(), rtol=2e-5, atol=1e-5)


def test_mv_vm_vv_against_torch():
    rng = np.random.default_rng(3)
    a, v, u = (rng.standard_normal(s).astype(F32) for s in ((5, 7), (7,), (5,)))
    ta, tv, tu = (torch.tensor(t, requires_grad=True) for t in (a, v, u))
    g = rng.standard_normal(5).astype(F32)
    (ta @ tv).backward(torch.tensor(g))
    da, dv = np.zeros_like(a), np.zeros_like(v)
    O.mv_backward(a, v, g, da, dv)
    assert np.allclose(O.mv_forward(a, v), a @ v, rtol=1e-5) and np.allclose(da, ta.grad.numpy(), rtol=1e-5)
    assert np.allclose(dv, tv.grad.numpy(), rtol=1e-5, atol=1e-6)
    ta.grad = None
    g2 = rng.standard_normal(7).astype(F32)
    (tu @ ta).backward(torch.tensor(g2))
    du, da = np.zeros_like(u), np.zeros_like(a)
    O.vm_backward(u, a, g2, du, da)
    assert np.allclose(du, tu.grad.numpy(), rtol=1e-5, atol=1e-6) and np.allclose(da, ta.grad.numpy(), rtol=1e-5)


def _torch_opt_run(make_opt, steps, w0, grads):
    tw = torch.tensor(w0.copy(), requires_grad=True)
    opt = make_opt([tw])
    for g in grads[:steps]:
        tw.grad = torch.tensor(g.copy())
        opt.step()
    return tw.detach().numpy()


def test_adam_family_against_torch_optim():
    rng = np.random.default_rng(11)
    w0 = rng.standard_normal(257).astype(F32)
    grads = [rng.standard_normal(257).astype(F32) for _ in range(4)]
    lam = 0.01
    # torch's weight_decay adds wd*w; the reference's L2 adds 2*lambda*w -> wd = 2*lambda
    for amsgrad in (False, True):
        w = w0.copy()
        m, v = np.zeros_like(w), np.zeros_like(w)
        vmax = np.zeros_like(w) if amsgrad else None
        for t, g in enumerate(grads, 1):
            O.adam_step(w, g.copy(), m, v, t, 1e-2, 0.9, 0.999, 1e-8, l2=lam, max_exp_avg_sq=vmax)
        want = _torch_opt_run(lambda p: torch.optim.Adam(p, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=2 * lam,
                                                         amsgrad=amsgrad), 4, w0, grads)
        assert np.allclose(w, want, rtol=2e-5, atol=2e-6), amsgrad
    for centered, mom in ((False, None), (True, None), (False, 0.9), (True, 0.9)):
        w = w0.copy()
        sq, ga, buf = np.zeros_like(w), np.zeros_like(w), np.zeros_like(w)
        for g in grads:
            O.rmsprop_step(w, g.copy(), sq, 1e-2, 0.99, 1e-8, momentum=mom, centered=centered, grad_avg=ga, buffer=buf,
                           l2=lam)
        want = _torch_opt_run(lambda p: torch.optim.RMSprop(p, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=2 * lam,
                                                            momentum=mom or 0.0, centered=centered), 4, w0, grads)
        assert np.allclose(w, want, rtol=5e-5, atol=5e-6), (centered, mom)
    w = w0.copy()
    s = np.zeros_like(w)
    for t, g in enumerate(grads, 1):
        O.adagrad_step(w, g.copy(), s, t, 1e-2, 0.1, 1e-10, l2=lam)
    want = _torch_opt_run(lambda p: torch.optim.Adagrad(p, lr=1e-2, lr_decay=0.1, eps=1e-10, weight_decay=2 * lam), 4, w0, grads)
    assert np.allclose(w, want, rtol=2e-5, atol=2e-6)


def test_penalties():
    w = np.array([-2.0, -0.0, 0.0, 3.0], F32)
    assert np.array_equal(O.penalize(w, l1=0.5), np.array([-0.5, -0.5, 0.5, 0.5], F32))    # f32::signum(+-0) = +-1
    assert np.array_equal(O.penalize(w, l2=0.25), 0.5 * w)
    assert np.allclose(O.penalize(w, l1=0.5, l2=0.25), O.penalize(w, l1=0.5) + O.penalize(w, l2=0.25))
