"""Algebra check of the v5 element math (spvipes_b200/csrc/nb_math.cuh) against scvi-tools' log_mixture_nb and its autograd
gradients, in float64 on the CPU.  Not part of the product path; run by hand after touching nb_math.cuh.

    python tools/check_nb_math.py
"""
import numpy as np
import torch

LOG2E, LN2, EPS = 1.4426950408889634, 0.6931471805599453, 1e-8


def scvi_log_mixture_nb(x, mu1, mu2, theta, pi_logits):
    lte = torch.log(theta + EPS)
    l1, l2 = torch.log(theta + mu1 + EPS), torch.log(theta + mu2 + EPS)
    lg = torch.lgamma(x + theta) - torch.lgamma(theta) - torch.lgamma(x + 1)
    nb1 = theta * (lte - l1) + x * (torch.log(mu1 + EPS) - l1) + lg
    nb2 = theta * (lte - l2) + x * (torch.log(mu2 + EPS) - l2) + lg
    lse = torch.logsumexp(torch.stack((nb1, nb2 - pi_logits)), dim=0)
    return lse - torch.nn.functional.softplus(-pi_logits)


def v5_forward(t, Ct, xp, xs, pi, th, exact):
    thE, Kc = th + EPS, th * np.log(th + EPS)
    rp, rs = 2.0 ** xp, 2.0 ** xs
    d1, d2 = rp + thE, rs + thE
    L1, L2 = np.log2(d1), np.log2(d2)
    x = t + th
    Lap, Las, fp, fs = xp, xs, 1.0, 1.0
    if exact:
        Lap, Las = np.log2(rp + EPS), np.log2(rs + EPS)
        fp, fs = rp / (rp + EPS), rs / (rs + EPS)
    m1, m2 = x * L1 - t * Lap, x * L2 - t * Las
    df = (m2 - m1) * LN2 + pi
    e, epi = 2.0 ** (-LOG2E * np.abs(df)), 2.0 ** (-LOG2E * np.abs(pi))
    o1, o2 = 1 + e, 1 + epi
    d12, oo = d1 * d2, o1 * o2
    r = 1.0 / (d12 * oo)
    rd, ro = r * oo, r * d12
    id1, id2, i1, i2 = rd * d2, rd * d1, ro * o2, ro * o1
    ll = (Kc + Ct) + (-LN2 * m1 + np.maximum(-df, 0)) - np.maximum(-pi, 0) + LN2 * np.log2(o1 * i2)
    wmin = e * i1
    wa = np.where(df >= 0, 1 - wmin, wmin)
    wb = 1 - wa
    ep = wa * (t * fp - x * id1 * rp)
    es = wb * (t * fs - x * id2 * rs)
    # backward pieces
    K1c = np.log(th + EPS) + th / (th + EPS)
    q1, q2 = x * id1, x * id2
    sneg = np.where(pi >= 0, epi, 1.0) * i2
    return ll, ep, es, sneg - wb, lambda Pt: (K1c + Pt) - wa * (L1 * LN2 + q1) - wb * (L2 * LN2 + q2)


def main():
    rng = np.random.default_rng(0)
    n = 200000
    c = rng.choice([0, 0, 0, 0, 1, 2, 3, 7, 15, 40, 300], n).astype(np.float64)
    t = np.log1p(c)
    th = np.exp(rng.normal(0, 1.5, n))
    yp, ys = rng.normal(-6, 4, n), rng.normal(-6, 4, n)   # natural-log rho
    pi = rng.normal(0, 3, n)
    T = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tt, tth, typ, tys, tpi = torch.tensor(t), T(th), T(yp), T(ys), T(pi)
    ref = scvi_log_mixture_nb(tt, torch.exp(typ), torch.exp(tys), tth, tpi)
    ref.sum().backward()
    Ct = (torch.lgamma(tt + tth) - torch.lgamma(tth) - torch.lgamma(tt + 1)).detach().numpy()
    Pt = (torch.digamma(tt + tth) - torch.digamma(tth)).detach().numpy()
    for exact in (True, False):
        sel = np.ones(n, bool) if exact else ~((t > 0) & (np.minimum(yp, ys) * LOG2E < -19.931568))
        ll, ep, es, dpi, dth = v5_forward(t, Ct, yp * LOG2E, ys * LOG2E, pi, th, exact)
        err = lambda a, b: float(np.max(np.abs(a - b)[sel] / (1 + np.abs(b)[sel])))
        print(f"exact={exact}: ll {err(ll, ref.detach().numpy()):.2e}  d/dyp {err(ep, typ.grad.numpy()):.2e}  d/dys {err(es, tys.grad.numpy()):.2e}"
              f"  d/dpi {err(dpi, tpi.grad.numpy()):.2e}  d/dtheta {err(dth(Pt), tth.grad.numpy()):.2e}")


if __name__ == "__main__":
    main()
