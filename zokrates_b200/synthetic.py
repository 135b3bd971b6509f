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


def make_layered(ctx, curve, n_constraints: int, seed: int = 0x5EED0003, distribution: str = "uniform", layers: int = 64):
    """Same family of circuits as `make`, generated layer by layer with vectorised numpy indexing and the
    witness values computed by batched field arithmetic on the GPU (`zkb_field_op`, the R1CS
    witness-arithmetic row a9/a11 of SURVEY.md §8) — seconds instead of minutes at 2^20 constraints.
    Operands of layer k come from layers < k.  Coefficients and inputs are uniform below 2^252."""
    c = _curve(curve)
    rng = np.random.RandomState((seed ^ 0xA5A5) & 0x7FFFFFFF)
    N = n_constraints
    ni, n_priv_in = 2, 6
    m0 = ni + n_priv_in
    bits = distribution == "bits"

    def rand_fr(k):
        v = rng.randint(0, 1 << 62, size=(k, 4), dtype=np.int64).astype(np.uint64) << np.uint64(2)
        v |= rng.randint(0, 4, size=(k, 4), dtype=np.int64).astype(np.uint64)
        v[:, 3] &= np.uint64((1 << 60) - 1)
        return v

    z = np.zeros((m0 + N, 4), dtype=np.uint64)
    z[0, 0] = 1
    if bits:
        z[1:m0, 0] = rng.randint(0, 2, size=m0 - 1).astype(np.uint64)
        z[1, 0] = 1
    else:
        z[1:m0] = rand_fr(m0 - 1)
    is_bit = np.zeros(m0 + N, dtype=bool)
    if bits:
        is_bit[1:m0] = True

    width = max(1, (N + layers - 1) // layers)
    A_cols, A_vals, A_cnt, B_cols, B_vals, B_cnt, C_cols, C_vals, C_cnt = [], [], [], [], [], [], [], [], []
    one = np.zeros((1, 4), dtype=np.uint64); one[0, 0] = 1
    two = one.copy(); two[0, 0] = 2
    minus1 = fr_array([c.r - 1])

    def mul(a, b):
        return ctx.field_op(0, 0, np.ascontiguousarray(a), np.ascontiguousarray(b)) if len(a) else a

    def add(a, b):
        return ctx.field_op(0, 1, np.ascontiguousarray(a), np.ascontiguousarray(b)) if len(a) else a

    row = 0
    while row < N:
        w = min(width, N - row)
        avail = m0 + row                      # variables defined so far
        out_idx = m0 + row + np.arange(w)
        lin = rng.rand(w) < (0.10 if bits else 0.25)
        k = np.where(lin, rng.randint(2, 5, size=w), 1)
        ia = (rng.rand(w, 4) * avail).astype(np.int64)
        ib = (rng.rand(w, 4) * avail).astype(np.int64)
        ca = rand_fr(w * 4).reshape(w, 4, 4)
        cb = rand_fr(w * 4).reshape(w, 4, 4)
        xor = np.zeros(w, dtype=bool)
        if bits:
            xor = ~lin
            bit_idx = np.flatnonzero(is_bit[:avail])
            ia[xor, 0] = bit_idx[(rng.rand(int(xor.sum())) * len(bit_idx)).astype(np.int64)]
            ib[xor, 0] = bit_idx[(rng.rand(int(xor.sum())) * len(bit_idx)).astype(np.int64)]
            ca[xor, 0] = two
            cb[xor, 0] = one
        plain = ~lin & ~xor
        ca[plain, 0] = one
        cb[plain, 0] = one
        tmask = np.arange(4)[None, :] < k[:, None]
        # values: av = sum_t ca_t * z[ia_t] over the lin rows; plain/xor rows are cheap
        val = np.zeros((w, 4), dtype=np.uint64)
        lr = np.flatnonzero(lin)
        if len(lr):
            av = np.zeros((len(lr), 4), dtype=np.uint64)
            bv = np.zeros((len(lr), 4), dtype=np.uint64)
            for t in range(4):
                sel = tmask[lr, t]
                if not sel.any():
                    continue
                rows_t = lr[sel]
                pa = mul(ca[rows_t, t], z[ia[rows_t, t]])
                pb = mul(cb[rows_t, t], z[ib[rows_t, t]])
                av[sel] = add(av[sel], pa)
                bv[sel] = add(bv[sel], pb)
            val[lr] = mul(av, bv)
        pr = np.flatnonzero(plain)
        if len(pr):
            val[pr] = mul(z[ia[pr, 0]], z[ib[pr, 0]])
        xr = np.flatnonzero(xor)
        if len(xr):
            val[xr, 0] = z[ia[xr, 0], 0] ^ z[ib[xr, 0], 0]
            is_bit[out_idx[xr]] = True
        z[out_idx] = val
        A_cols.append(ia[tmask]); A_vals.append(ca[tmask]); A_cnt.append(k)
        B_cols.append(ib[tmask]); B_vals.append(cb[tmask]); B_cnt.append(k)
        # C row: [w] for product / lin rows; [a, b, -w] for xor rows
        ck = np.where(xor, 3, 1)
        cc = np.zeros((w, 3), dtype=np.int64)
        cv = np.zeros((w, 3, 4), dtype=np.uint64)
        cc[:, 0] = np.where(xor, ia[:, 0], out_idx); cv[:, 0] = one
        cc[:, 1] = ib[:, 0]; cv[:, 1] = one
        cc[:, 2] = out_idx; cv[:, 2] = minus1
        cmask = np.arange(3)[None, :] < ck[:, None]
        C_cols.append(cc[cmask]); C_vals.append(cv[cmask]); C_cnt.append(ck)
        row += w

    def csr(cols, vals, cnts):
        cnt = np.concatenate(cnts).astype(np.uint64)
        rowptr = np.zeros(N + 1, dtype=np.uint64)
        np.cumsum(cnt, out=rowptr[1:])
        return rowptr, np.concatenate(cols).astype(np.uint32), np.concatenate(vals).astype(np.uint64)

    r1cs = R1CS(c.name, N, ni, n_priv_in + N, csr(A_cols, A_vals, A_cnt), csr(B_cols, B_vals, B_cnt),
                csr(C_cols, C_vals, C_cnt))
    return r1cs, z
