"""Generate tests/golden/known_answers.json with mpmath at 50 digits.

The reference (Filippo-Galli/Split_and_merge_Gibbs_sampling) ships no tests and
no golden vectors and cannot be built in this image (R/Rcpp/GSL absent), so the
oracle is pinned against closed-form definitions evaluated in high precision:

  dhamming(x,c,s,m)      code/common_functions.cpp:355-377
  norm_const2(d,c,m)     code/hyperg.cpp:11-48      (log(d+1)+(d+c)log m - log 2F1(d+c,1;d+2;(m-1)/m))
  logdensity_hig         code/split_merge.cpp:6-18
  lF_conK2(u,d,c,m,lK)   code/hyperg.cpp:183-217
  HIG inverse CDF in u   code/hyperg.cpp:221-287    (root of lF_conK2(u) = log Omega)
  rhig switching rule    code/hyperg.cpp:359        (qbeta(0.1,w+1,v-1) < (m-1)/m)

Run:  python tests/golden/make_known_answers.py
"""
import json
import os

import mpmath as mp

mp.mp.dps = 50


def dhamming(x, c, s, m):
    s = mp.mpf(s)
    return -(0 if x == c else 1) / s - mp.log(1 + (m - 1) / mp.exp(1 / s))


def f21(a, b, c, x):
    return mp.hyp2f1(a, b, c, x, maxterms=10**7)


def ibeta(x, a, b):
    """Regularised incomplete beta I_x(a,b) through the all-positive series
    I_x = x^a (1-x)^b / (a B(a,b)) * 2F1(a+b, 1; a+1; x), reflected about the mean."""
    x, a, b = mp.mpf(x), mp.mpf(a), mp.mpf(b)
    if x > (a + 1) / (a + b + 2):
        return 1 - ibeta(1 - x, b, a)
    return mp.exp(a * mp.log(x) + b * mp.log(1 - x) - mp.log(a) - mp.log(mp.beta(a, b))) * f21(a + b, 1, a + 1, x)


def norm_const2(d, c, m):
    d, c, m = mp.mpf(d), mp.mpf(c), mp.mpf(m)
    return mp.log(d + 1) + (d + c) * mp.log(m) - mp.log(f21(d + c, 1, d + 2, (m - 1) / m))


def logdensity_hig(s, v, w, m):
    s, v, w, m = map(mp.mpf, (s, v, w, m))
    return norm_const2(w, v, m) - (v + w) * mp.log(1 + mp.exp(-1 / s) * (m - 1)) - (w + 1) / s - 2 * mp.log(s)


def lF(u, d, c, m):
    u, d, c, m = map(mp.mpf, (u, d, c, m))
    lK = norm_const2(d, c, m)
    x = u * (m - 1) / (1 + u * (m - 1))
    return lK - mp.log(d + 1) + (d + 1) * mp.log(u) - (d + c) * mp.log(1 + u * (m - 1)) + mp.log(f21(1, d + c, d + 2, x))


def hig_inv_u(omega, v, w, m):
    """u with CDF(u) = omega, CDF in u-space = I_x(w+1, v-1)/I_{(m-1)/m}(w+1, v-1)."""
    v, w, m, omega = map(mp.mpf, (v, w, m, omega))
    a, b = w + 1, v - 1
    tot = ibeta((m - 1) / m, a, b)
    f = lambda x: ibeta(x, a, b) - omega * tot
    lo, hi = mp.mpf(0), (m - 1) / m
    for _ in range(200):
        mid = (lo + hi) / 2
        if f(mid) < 0:
            lo = mid
        else:
            hi = mid
    x = (lo + hi) / 2
    return x / ((m - 1) * (1 - x))


out = {"dhamming": [], "hyp2f1": [], "norm_const2": [], "logdensity_hig": [], "lF_conK2": [], "pbeta": [],
       "hig_inv_u": [], "rhig_branch": []}

for (x, c, s, m) in [(1, 1, 0.5, 2), (1, 2, 0.5, 2), (3, 3, 1.25, 6), (3, 5, 1.25, 6), (2, 2, 0.05, 5), (4, 1, 7.5, 5),
                     (1, 1, 0.31, 4), (2, 3, 0.31, 4)]:
    out["dhamming"].append({"x": x, "c": c, "s": s, "m": m, "val": float(dhamming(x, c, s, m))})

for (a, b, c, x) in [(6.25, 1, 2.25, 0.5), (3.5, 1, 2.5, 5 / 6), (3.5, 1, 2.5, 0.5), (106.25, 1, 32.25, 0.5),
                     (503.5, 1, 152.5, 0.8), (1, 6.25, 2.25, 0.2), (1, 106.25, 32.25, 0.31), (40.0, 1, 12.0, 0.75)]:
    out["hyp2f1"].append({"a": a, "b": b, "c": c, "x": x, "val": float(f21(a, b, c, mp.mpf(x)))})

for (d, c, m) in [(0.25, 6, 2), (0.5, 3, 6), (30.25, 76, 2), (150.5, 353, 5), (0.25, 6, 5), (10.25, 26, 4),
                  (700.25, 1306, 5), (3500.25, 6506, 5), (2.5, 1.5, 3)]:
    out["norm_const2"].append({"d": d, "c": c, "m": m, "val": float(norm_const2(d, c, m))})

for (s, v, w, m) in [(0.7, 6, 0.25, 2), (1.3, 3, 0.5, 6), (0.2, 26, 30.25, 2), (0.5, 6, 0.25, 5), (0.45, 1306, 700.25, 5),
                     (2.5, 6, 0.25, 4)]:
    out["logdensity_hig"].append({"s": s, "v": v, "w": w, "m": m, "val": float(logdensity_hig(s, v, w, m))})

for (u, d, c, m) in [(0.5, 0.25, 6, 2), (0.1, 0.25, 6, 5), (0.9, 0.5, 3, 6), (0.3, 30.25, 76, 2), (0.05, 150.5, 353, 5)]:
    out["lF_conK2"].append({"u": u, "d": d, "c": c, "m": m, "lK": float(norm_const2(d, c, m)), "val": float(lF(u, d, c, m))})

for (x, a, b) in [(0.5, 1.25, 5), (0.8, 1.25, 5), (0.5, 19.25, 7), (0.5, 16.25, 10), (0.8, 701.25, 1305), (0.8, 3501.25, 6505),
                  (0.75, 30.5, 9.5), (0.5, 31.25, 75), (0.5, 75, 31.25)]:
    val = ibeta(x, a, b)
    out["pbeta"].append({"x": x, "a": a, "b": b, "val": float(val), "logval": float(mp.log(val))})

for (omega, v, w, m) in [(0.5, 6, 0.25, 2), (0.1, 6, 0.25, 5), (0.93, 3, 0.5, 6), (0.37, 26, 30.25, 2), (0.62, 353, 150.5, 5),
                         (0.999, 6, 0.25, 4), (0.001, 6, 0.25, 4), (0.5, 8, 18.25, 2), (0.25, 1306, 700.25, 5)]:
    out["hig_inv_u"].append({"omega": omega, "v": v, "w": w, "m": m, "u": float(hig_inv_u(omega, v, w, m))})

# switching rule: Beta branch iff pbeta((m-1)/m; w+1, v-1) > 0.1  (and v > 1)
for (v, w, m) in [(6, 0.25, 2), (6, 0.25, 5), (3, 0.5, 6), (8, 18.25, 2), (11, 15.25, 2), (0.8, 0.25, 2), (26, 30.25, 2),
                  (353, 150.5, 5), (10, 40.25, 2)]:
    a, b = w + 1, v - 1
    if b <= 0:
        br = 0
        pv = float("nan")
    else:
        pv = float(ibeta(mp.mpf(m - 1) / m, a, b))
        br = int(pv > 0.1)
    out["rhig_branch"].append({"v": v, "w": w, "m": m, "pbeta": pv, "beta_branch": br})

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
