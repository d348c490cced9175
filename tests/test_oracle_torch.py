"""Second pin of the oracle (SURVEY.md 8-c "oracle form"): torch-CPU autograd on random shapes.

The reference's golden vectors (tests/test_oracle_goldens.py) use small, mostly uniform tensors; they cannot see the
two places where the oracle implements the INTENDED maths rather than the reference's defective code (un-broadcast of
non-symmetric gradients, `utils.rs:152-192`; the column layout of `convolution_backward_input`,
`convolution/mod.rs:146-189`).  Random kernels, gradients and shapes against an independent autograd close that gap.
CPU only; torch is the checker here, never part of the product path."""
import numpy as np
import pytest

import oracle as O

torch = pytest.importorskip("torch")
F32 = np.float32


def T(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=F32)).double()
    t.requires_grad_(grad)
    return t


def close(got, want, tol=2e-5):
    want = want.detach().numpy() if hasattr(want, "detach") else want
    scale = float(np.sqrt((want.astype(np.float64) ** 2).mean())) + 1e-12
    return bool(np.all(np.abs(got.astype(np.float64) - want) <= tol * (scale + np.abs(want))))


@pytest.mark.parametrize("nd,stride,dil,groups", [(2, (1, 1), (1, 1), 1), (2, (2, 1), (1, 2), 1), (2, (1, 2), (2, 1), 2),
                                                  (2, (3, 2), (1, 1), 4), (1, (2,), (2,), 1), (3, (1, 2, 1), (1, 1, 2), 2)])
def test_convolution_matches_autograd(nd, stride, dil, groups):
    rng = np.random.default_rng(10 * nd + groups + stride[0])
    cin, cout = 4 * groups if groups > 1 else 3, 4 * groups if groups > 1 else 5
    spatial = {1: (19,), 2: (11, 13), 3: (6, 9, 7)}[nd]
    ksz = {1: (3,), 2: (3, 2), 3: (2, 3, 2)}[nd]
    x = rng.uniform(-1, 1, (3, cin) + spatial).astype(F32)
    w = rng.uniform(-1, 1, (cout, cin // groups) + ksz).astype(F32)
    y = O.conv_forward(x, w, stride, dil, groups)
    g = rng.uniform(-1, 1, y.shape).astype(F32)
    dx, dw = np.zeros_like(x), np.zeros_like(w)
    O.conv_backward_input(dx, g, w, stride, dil, groups)
    O.conv_backward_kernel(dw, g, x, stride, dil, groups)
    tx, tw = T(x, True), T(w, True)
    fn = {1: torch.nn.functional.conv1d, 2: torch.nn.functional.conv2d, 3: torch.nn.functional.conv3d}[nd]
    ty = fn(tx, tw, None, stride, 0, dil, groups)
    ty.backward(T(g))
    assert y.shape == tuple(ty.shape)
    assert close(y, ty) and close(dx, tx.grad) and close(dw, tw.grad)
    # accumulate protocol (every backward += into the operand gradient, vardiff.rs:125-141)
    O.conv_backward_input(dx, g, w, stride, dil, groups)
    assert close(dx, 2 * tx.grad)


@pytest.mark.parametrize("shape_l,shape_r", [((7, 5), (5,)), ((6, 4, 3, 5), (4, 1, 1)), ((3, 1, 5), (4, 1)),
                                             ((5,), (2, 3, 5)), ((2, 3), (2, 3))])
def test_broadcast_add_and_unbroadcast_match_autograd(shape_l, shape_r):
    rng = np.random.default_rng(len(shape_l) * 7 + len(shape_r))
    l, r = rng.uniform(-1, 1, shape_l).astype(F32), rng.uniform(-1, 1, shape_r).astype(F32)
    y = O.add_forward(l, r)
    g = rng.uniform(-1, 1, y.shape).astype(F32)           # NON-symmetric gradient: what the reference's tests lack
    dl, dr = np.zeros_like(l), np.zeros_like(r)
    O.add_backward(g, dl, dr)
    tl, tr = T(l, True), T(r, True)
    ty = tl + tr
    ty.backward(T(g))
    assert y.shape == tuple(ty.shape) == O.cobroadcast(shape_l, shape_r)
    assert close(y, ty) and close(dl, tl.grad) and close(dr, tr.grad)


@pytest.mark.parametrize("final,loss", [("softmax", "mse"), ("log_softmax", "nll"), ("log_softmax", "mse")])
def test_mlp_head_matches_autograd(final, loss):
    """Linear -> ReLU -> Linear -> (log-)softmax -> loss: forward value and every parameter / input gradient"""
    rng = np.random.default_rng(5)
    n, i, h, c = 9, 6, 8, 4
    x = rng.uniform(-1, 1, (n, i)).astype(F32)
    w1, b1 = O.uniform_init(rng, (h, i), 0.4), O.uniform_init(rng, (h,), 0.4)
    w2, b2 = O.uniform_init(rng, (c, h), 0.4), O.uniform_init(rng, (c,), 0.4)
    cls = rng.integers(0, c, n)
    tgt = np.eye(c, dtype=F32)[cls]
    # oracle forward
    z1 = O.linear_forward(x, w1, b1)
    a1 = O.relu_forward(z1)
    z2 = O.linear_forward(a1, w2, b2)
    p = O.softmax_forward(z2, 1) if final == "softmax" else O.log_softmax_forward(z2, 1)
    if loss == "mse":
        lv = O.mse_forward(p, tgt, "mean")
        dp = np.zeros_like(p)
        O.mse_backward(p, tgt, F32(1.0), dp, "mean")
    else:
        lv = O.nll_forward(p, cls.astype(F32), "mean")
        dp = np.zeros_like(p)
        O.nll_backward(cls.astype(F32), F32(1.0), dp, "mean")
    dz2 = np.zeros_like(z2)
    (O.softmax_backward if final == "softmax" else O.log_softmax_backward)(p, dp, dz2, 1)
    da1, dw2, db2 = np.zeros_like(a1), np.zeros_like(w2), np.zeros_like(b2)
    O.linear_backward(a1, w2, dz2, da1, dw2, db2)
    dz1 = np.zeros_like(z1)
    O.relu_backward(z1, da1, dz1)
    dx, dw1, db1 = np.zeros_like(x), np.zeros_like(w1), np.zeros_like(b1)
    O.linear_backward(x, w1, dz1, dx, dw1, db1)
    # torch
    tx, tw1, tb1, tw2, tb2 = T(x, True), T(w1, True), T(b1, True), T(w2, True), T(b2, True)
    tz2 = torch.relu(tx @ tw1.T + tb1) @ tw2.T + tb2
    tp = torch.softmax(tz2, 1) if final == "softmax" else torch.log_softmax(tz2, 1)
    tl = ((tp - T(tgt)) ** 2).mean() if loss == "mse" else torch.nn.functional.nll_loss(tp, torch.from_numpy(cls))
    tl.backward()
    assert abs(float(lv) - float(tl.detach())) <= 1e-5 * (1 + abs(float(tl.detach())))
    for got, want in ((dx, tx.grad), (dw1, tw1.grad), (db1, tb1.grad), (dw2, tw2.grad), (db2, tb2.grad)):
        assert close(got, want, tol=5e-5)


def test_pad_sum_mean_match_autograd():
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, (2, 3, 5, 4)).astype(F32)
    y = O.pad_forward(x, (2, 1), 0.5)
    tx = T(x, True)
    ty = torch.nn.functional.pad(tx, (1, 1, 2, 2), value=0.5)
    assert np.array_equal(y, ty.detach().numpy().astype(F32))           # copies: bit exact
    g = rng.uniform(-1, 1, y.shape).astype(F32)
    dx = np.zeros_like(x)
    O.pad_backward(g, dx, (2, 1))
    ty.backward(T(g))
    assert np.array_equal(dx, tx.grad.numpy().astype(F32))
    for fwd, bwd, tf in ((O.sum_forward, O.sum_backward, torch.sum), (O.mean_forward, O.mean_backward, torch.mean)):
        tx = T(x, True)
        tv = tf(tx)
        tv.backward(torch.tensor(0.7, dtype=torch.float64))
        d = np.zeros_like(x)
        bwd(F32(0.7), d)
        assert abs(float(fwd(x)) - float(tv.detach())) <= 1e-5 * (1 + abs(float(tv.detach()))) and close(d, tx.grad)


@pytest.mark.parametrize("momentum,nesterov", [(None, False), (0.9, False), (0.9, True)])
def test_sgd_matches_torch_optimizer(momentum, nesterov):
    """SGDParam::optimize (sgd/mod.rs:191-231) with the L2 penalty `g += 2*lambda*w` (penalty.rs:63-67) is torch's SGD
    with weight_decay = 2*lambda (dampening 0), three steps with a state buffer"""
    rng = np.random.default_rng(3)
    w = rng.uniform(-1, 1, (5, 4)).astype(F32)
    tw = T(w, True)
    lam = 0.05
    opt = torch.optim.SGD([tw], lr=0.1, momentum=momentum or 0.0, nesterov=nesterov, weight_decay=2 * lam)
    buf = None
    for step in range(3):
        g = rng.uniform(-1, 1, w.shape).astype(F32)
        tw.grad = T(g)
        opt.step()
        buf = O.sgd_step(w, g.copy(), 0.1, lam, momentum, 0.0 if momentum else None, nesterov, buf)
        assert close(w, tw, tol=2e-5), step


def test_sgd_dampening_follows_the_reference_not_torch():
    """With dampening the reference scales the FIRST gradient too (its buffer starts at zero and every step does
    buf = mu*buf + (1-d)*g, sgd/mod.rs:209-219); torch seeds the buffer with the undamped first gradient.  The oracle
    follows the reference: closed form over three steps."""
    rng = np.random.default_rng(4)
    w0 = rng.uniform(-1, 1, (6,)).astype(F32)
    gs = [rng.uniform(-1, 1, w0.shape).astype(F32) for _ in range(3)]
    mu, d, lr = 0.8, 0.1, 0.1
    w, buf = w0.copy(), None
    for g in gs:
        buf = O.sgd_step(w, g.copy(), lr, 0.0, mu, d, False, buf)
    b, want = np.zeros(6), w0.astype(np.float64)
    for g in gs:
        b = mu * b + (1 - d) * g
        want = want - lr * b
    assert np.allclose(w, want, rtol=1e-5, atol=1e-6)
