"""Synthetic R1CS generator for the benchmark configurations.

Constraint i is  <A_i, z> * <B_i, z> = z[w_i]  with w_i a fresh witness variable, so the assignment is
computable row by row.  75 % of the rows are plain products z[u] * z[v]; 25 % carry 2-4 term linear
combinations with uniform Fr coefficients.  One public input (l = 1) so that N = 2^k - 2 gives a
domain of exactly 2^k.  Two witness distributions: "uniform" (full-width values: MSM worst case) and
"bits" (about 90 % of the values are 0/1, like a SHA-256 circuit).
Deterministic in (seed, log_n, distribution).
"""
from __future__ import annotations

import numpy as np

from ._lib import fr_array
from .curves import curve as _curve
from .r1cs import R1CS


def make(curve, n_constraints: int, seed: int = 0x5EED0003, distribution: str = "uniform"):
    c = _curve(curve)
    r = c.r
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    N = n_constraints
    ni = 2                      # one + one public input
    n_priv_in = 6
    m0 = ni + n_priv_in

    def rand_fr(k):
        # uniform-ish field elements from 5 x 52-bit draws (python ints)
        parts = rng.randint(0, 1 << 52, size=(k, 5), dtype=np.int64)
        return [(int(p[0]) | (int(p[1]) << 52) | (int(p[2]) << 104) | (int(p[3]) << 156) | (int(p[4]) << 208)) % r
                for p in parts]

    bits = distribution == "bits"
    z = [1] + (rand_fr(m0 - 1) if not bits else [int(v) for v in rng.randint(0, 2, size=m0 - 1)])
    if bits:
        z[1] = 1
    lin_rows = rng.rand(N) < (0.25 if not bits else 0.10)
    nterms = rng.randint(2, 5, size=N)
    # operands are drawn from variables that already exist when row i is evaluated: index < m0 + i
    ua = (rng.rand(N, 4) * (m0 + np.arange(N))[:, None]).astype(np.int64)
    ub = (rng.rand(N, 4) * (m0 + np.arange(N))[:, None]).astype(np.int64)
    n_lin = int(lin_rows.sum())
    coef = rand_fr(2 * 4 * n_lin) if n_lin else []
    a_rowptr = np.zeros(N + 1, dtype=np.uint64); b_rowptr = np.zeros(N + 1, dtype=np.uint64)
    c_rowptr = np.zeros(N + 1, dtype=np.uint64)
    a_col, a_val, b_col, b_val, c_col, c_val = [], [], [], [], [], []
    bit_vars = list(range(1, m0)) if bits else []
    frac = rng.rand(N, 2)
    ci = 0
    for i in range(N):
        w = m0 + i
        if lin_rows[i]:
            k = int(nterms[i])
            ca = coef[ci:ci + k]; cb = coef[ci + 4:ci + 4 + k]; ci += 8
            av = 0; bv = 0
            for t in range(k):
                ja, jb = int(ua[i, t]), int(ub[i, t])
                a_col.append(ja); a_val.append(ca[t]); av += ca[t] * z[ja]
                b_col.append(jb); b_val.append(cb[t]); bv += cb[t] * z[jb]
            z.append(av % r * (bv % r) % r)
            c_col.append(w); c_val.append(1)
        elif bits:
            # XOR gate on two bit variables: (2a) * b = a + b - w   (w = a xor b stays balanced)
            ja = bit_vars[int(frac[i, 0] * len(bit_vars))]; jb = bit_vars[int(frac[i, 1] * len(bit_vars))]
            a_col.append(ja); a_val.append(2); b_col.append(jb); b_val.append(1)
            c_col += [ja, jb, w]; c_val += [1, 1, r - 1]
            z.append(z[ja] ^ z[jb])
            bit_vars.append(w)
        else:
            ja, jb = int(ua[i, 0]), int(ub[i, 0])
            a_col.append(ja); a_val.append(1); b_col.append(jb); b_val.append(1)
            z.append(z[ja] * z[jb] % r)
            c_col.append(w); c_val.append(1)
        a_rowptr[i + 1] = len(a_col); b_rowptr[i + 1] = len(b_col); c_rowptr[i + 1] = len(c_col)
    c_col = np.array(c_col, dtype=np.uint32)
    c_val = fr_array(c_val)
    nw = n_priv_in + N
    r1cs = R1CS(c.name, N, ni, nw,
                (a_rowptr, np.array(a_col, dtype=np.uint32), fr_array(a_val)),
                (b_rowptr, np.array(b_col, dtype=np.uint32), fr_array(b_val)),
                (c_rowptr, c_col, c_val))
    return r1cs, fr_array(z)
