"""TEST INFRASTRUCTURE — CPU restatement of the GM17 path, python big integers.

Restates what `impl Backend<T, GM17> for Ark` reaches (/root/reference/zokrates_ark/src/gm17.rs:19-75):
`ark_gm17::GM17::circuit_specific_setup`, `GM17::prove` (= `create_random_proof`: d1, d2, r drawn in that order, then
`create_proof`), `verify_proof`, and the square-arithmetic-program reduction `R1CStoSAP` they share.  ark-gm17 0.3.0 is an
external crate (Cargo.lock pins it next to ark-groth16; its sources are NOT under /root/reference), so this file is written
from the published algorithm — Groth & Maller, "Snarks of Knowledge from Square Arithmetic Programs" (CRYPTO 2017), Fig. 4 —
and from the structure of the crate (field order of `ProvingKey` / `VerifyingKey`, the R1CS -> SAP embedding with two square
constraints per R1CS row and two per public input, the d1 / d2 masking of a(x) and c(x)).

PARITY UNPINNED: the reference holds no GM17 golden proof or key (its tests assert `verify()` only, gm17.rs:113-160) and the
crate cannot be built here, so nothing in this file has been compared with real ark-gm17 output.  What IS checked
(tests/test_gm17.py): the proofs satisfy both GM17 pairing equations under the verifying key, they equal the trapdoor
prediction computed with Fr arithmetic alone, and the GPU prover (zkb_gm17_prove) produces the same bytes.

SAP embedding (R1CStoSAP): R1CS row i, <A_i,z><B_i,z> = <C_i,z>, becomes
    (A_i + B_i)^2 = 4 C_i + x_i      and      (A_i - B_i)^2 = x_i          with one extra variable x_i = (A_i - B_i)^2,
the constant row 1^2 = 1, and for every public input j >= 1
    (z_j + 1)^2 = 4 z_j + y_j        and      (z_j - 1)^2 = y_j            with y_j = (z_j - 1)^2,
so the SAP has 2N + 2(l-1) + 1 rows (domain = next power of two) and (m - 1) + N + (l - 1) variables besides the constant.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

from .ark import Domain, R1CS, de_g1, de_g2, eval_row, ser_g1, ser_g2
from .ff import CurveParams, g1_group, g2_group, inv_mod, pairing_product_is_one


@dataclass
class Gm17Trapdoor:
    alpha: int
    beta: int
    gamma: int
    tau: int
    g1_k: int = 1       # g = g1_k * (standard G1 generator); ark draws a random generator
    g2_k: int = 1


@dataclass
class Gm17ProvingKey:
    # vk
    h_g2: tuple
    g_alpha_g1: tuple
    h_beta_g2: tuple
    g_gamma_g1: tuple
    h_gamma_g2: tuple
    query: list
    # pk
    a_query: list
    b_query: list
    c_query_1: list
    c_query_2: list
    g_gamma_z: tuple
    h_gamma_z: tuple
    g_ab_gamma_z: tuple
    g_gamma2_z2: tuple
    g_gamma2_z_t: list


def sap_sizes(r1cs: R1CS):
    N, ni, m = r1cs.num_constraints, r1cs.num_instance, r1cs.num_variables
    rows = 2 * N + 2 * (ni - 1) + 1
    sap_vars = m + N + (ni - 1)            # columns incl. the constant: [z | x_0..x_{N-1} | y_1..y_{l-1}]
    return rows, sap_vars


def sap_at_tau(c: CurveParams, r1cs: R1CS, tau: int):
    """R1CStoSAP::instance_map_with_evaluation: u_i(tau) ("a"), w_i(tau) ("c") per SAP variable, Z(tau), the domain."""
    q = c.r
    N, ni, m = r1cs.num_constraints, r1cs.num_instance, r1cs.num_variables
    rows, nv = sap_sizes(r1cs)
    d = Domain(c, rows)
    u = d.lagrange_at(tau)
    a, cc = [0] * nv, [0] * nv
    x_off, y_off, e_off = m, m + N - 1, 2 * N        # y_j sits at y_off + j (j >= 1); extra rows start at e_off
    for i in range(N):
        u_add, u_sub = (u[2 * i] + u[2 * i + 1]) % q, (u[2 * i] - u[2 * i + 1]) % q
        for col, k in r1cs.a[i]:
            a[col] = (a[col] + u_add * k) % q
        for col, k in r1cs.b[i]:
            a[col] = (a[col] + u_sub * k) % q
        for col, k in r1cs.c[i]:
            cc[col] = (cc[col] + 4 * u[2 * i] * k) % q
        cc[x_off + i] = (cc[x_off + i] + u[2 * i] + u[2 * i + 1]) % q
    a[0] = (a[0] + u[e_off]) % q
    cc[0] = (cc[0] + u[e_off]) % q
    for j in range(1, ni):
        u1, u2 = u[e_off + 2 * j - 1], u[e_off + 2 * j]
        a[j] = (a[j] + u1 + u2) % q
        a[0] = (a[0] + u1 - u2) % q
        cc[j] = (cc[j] + 4 * u1) % q
        cc[y_off + j] = (cc[y_off + j] + u1 + u2) % q
    zt = (pow(tau, d.n, q) - 1) % q
    return d, a, cc, zt


def sap_assignment(c: CurveParams, r1cs: R1CS, z: List[int]) -> List[int]:
    """full_input_assignment of R1CStoSAP::witness_map: z, then x_i = (A_i - B_i)^2, then y_j = (z_j - 1)^2."""
    q = c.r
    xs = [pow((eval_row(c, r1cs.a[i], z) - eval_row(c, r1cs.b[i], z)) % q, 2, q) for i in range(r1cs.num_constraints)]
    ys = [pow((z[j] - 1) % q, 2, q) for j in range(1, r1cs.num_instance)]
    return [v % q for v in z] + xs + ys


def setup(c: CurveParams, r1cs: R1CS, td: Gm17Trapdoor) -> Gm17ProvingKey:
    """ark-gm17 generate_parameters with an explicit trapdoor (the reference draws alpha, beta, gamma, g, h, t from its rng)."""
    G1, G2 = g1_group(c), g2_group(c)
    q = c.r
    d, a, cc, zt = sap_at_tau(c, r1cs, td.tau)
    ni = r1cs.num_instance
    nv = len(a)
    g, h = G1.mul(c.g1, td.g1_k), G2.mul(c.g2, td.g2_k)
    ab = (td.alpha + td.beta) % q
    g2z = td.gamma * td.gamma % q * zt % q
    tp, powers = 1, []
    for _ in range(d.n + 1):
        powers.append(g2z * tp % q)
        tp = tp * td.tau % q
    return Gm17ProvingKey(
        h_g2=h, g_alpha_g1=G1.mul(g, td.alpha), h_beta_g2=G2.mul(h, td.beta), g_gamma_g1=G1.mul(g, td.gamma),
        h_gamma_g2=G2.mul(h, td.gamma),
        query=[G1.mul(g, (td.gamma * cc[i] + ab * a[i]) % q) for i in range(ni)],
        a_query=[G1.mul(g, a[i] * td.gamma % q) for i in range(nv)],
        b_query=[G2.mul(h, a[i] * td.gamma % q) for i in range(nv)],
        c_query_1=[G1.mul(g, (td.gamma * td.gamma % q * cc[i] + ab * td.gamma % q * a[i]) % q) for i in range(ni, nv)],
        c_query_2=[G1.mul(g, 2 * g2z * a[i] % q) for i in range(nv)],
        g_gamma_z=G1.mul(g, td.gamma * zt % q), h_gamma_z=G2.mul(h, td.gamma * zt % q),
        g_ab_gamma_z=G1.mul(g, ab * td.gamma % q * zt % q), g_gamma2_z2=G1.mul(g, g2z * zt % q),
        g_gamma2_z_t=[G1.mul(g, s) for s in powers])


def witness_map(c: CurveParams, r1cs: R1CS, z: List[int], d1: int, d2: int):
    """R1CStoSAP::witness_map: (full assignment, h coefficients (n + 1 of them))."""
    q = c.r
    N, ni = r1cs.num_constraints, r1cs.num_instance
    rows, nv = sap_sizes(r1cs)
    d = Domain(c, rows)
    n = d.n
    full = sap_assignment(c, r1cs, z)
    m = r1cs.num_variables
    a = [0] * n
    cv = [0] * n
    for i in range(N):
        az, bz, cz = eval_row(c, r1cs.a[i], z), eval_row(c, r1cs.b[i], z), eval_row(c, r1cs.c[i], z)
        a[2 * i], a[2 * i + 1] = (az + bz) % q, (az - bz) % q
        cv[2 * i], cv[2 * i + 1] = (4 * cz + full[m + i]) % q, full[m + i]
    a[2 * N] = 1
    cv[2 * N] = 1
    for j in range(1, ni):
        a[2 * N + 2 * j - 1], a[2 * N + 2 * j] = (z[j] + 1) % q, (z[j] - 1) % q
        y = full[m + N - 1 + j]
        cv[2 * N + 2 * j - 1], cv[2 * N + 2 * j] = (4 * z[j] + y) % q, y
    a_coeff = d.ifft(a)
    h = [2 * d1 * x % q for x in a_coeff]
    h[0] = (h[0] - d2 - d1 * d1) % q
    h.append(d1 * d1 % q)
    a_cos = d.coset_fft(a_coeff)
    c_cos = d.coset_fft(d.ifft(cv))
    zinv = inv_mod((pow(d.g, n, q) - 1) % q, q)
    quot = d.coset_ifft([(x * x - y) % q * zinv % q for x, y in zip(a_cos, c_cos)])
    for i in range(n - 1):
        h[i] = (h[i] + quot[i]) % q
    return full, h


def prove(c: CurveParams, pk: Gm17ProvingKey, r1cs: R1CS, z: List[int], d1: int, d2: int, r: int):
    """ark-gm17 create_proof.  Returns affine (A, B, C)."""
    G1, G2 = g1_group(c), g2_group(c)
    q = c.r
    ni = r1cs.num_instance
    full, h = witness_map(c, r1cs, z, d1, d2)
    rest = full[1:]
    aux = full[ni:]
    add1 = G1.add
    g_a = add1(add1(add1(G1.mul(pk.g_gamma_z, r), pk.a_query[0]), G1.mul(pk.g_gamma_z, d1)), G1.msm_naive(pk.a_query[1:], rest))
    g_b = G2.add(G2.add(G2.add(G2.mul(pk.h_gamma_z, r), pk.b_query[0]), G2.mul(pk.h_gamma_z, d1)), G2.msm_naive(pk.b_query[1:], rest))
    c2 = add1(pk.c_query_2[0], G1.msm_naive(pk.c_query_2[1:], rest))
    g_c = G1.msm_naive(pk.c_query_1, aux)
    g_c = add1(g_c, G1.mul(pk.g_gamma2_z2, r * r % q))
    g_c = add1(g_c, G1.mul(pk.g_ab_gamma_z, (r + d1) % q))
    g_c = add1(g_c, G1.mul(pk.g_gamma2_z2, 2 * r * d1 % q))
    g_c = add1(g_c, G1.mul(c2, r))
    g_c = add1(g_c, G1.mul(pk.g_gamma2_z_t[0], d2))
    g_c = add1(g_c, G1.msm_naive(pk.g_gamma2_z_t, h))
    return g_a, g_b, g_c


def trapdoor_expected_proof(c: CurveParams, r1cs: R1CS, td: Gm17Trapdoor, z: List[int], d1: int, d2: int, r: int):
    """(A, B, C) from the trapdoor with Fr arithmetic and one scalar multiplication each (Groth-Maller Fig. 4 with a(x) masked by
    d1 Z and c(x) by d2 Z) — no FFT, no MSM, no key."""
    G1, G2 = g1_group(c), g2_group(c)
    q = c.r
    d, a, cc, zt = sap_at_tau(c, r1cs, td.tau)
    full = sap_assignment(c, r1cs, z)
    ni = r1cs.num_instance
    a0 = sum(x * y for x, y in zip(a, full)) % q
    c0 = sum(x * y for x, y in zip(cc, full)) % q
    at = (a0 + d1 * zt) % q                               # a(tau), c(tau) of the masked polynomials
    ct = (c0 + d2 * zt) % q
    ht = (at * at - ct) % q * inv_mod(zt, q) % q          # h(tau)
    ga, ab = td.gamma, (td.alpha + td.beta) % q
    a_dlog = ga * (at + r * zt) % q
    aux = sum((ga * ga % q * cc[i] + ab * ga % q * a[i]) * full[i] for i in range(ni, len(full))) % q
    c_dlog = (aux + d1 * ab % q * ga % q * zt + d2 * ga % q * ga % q * zt
              + r * r % q * ga % q * ga % q * zt % q * zt + r * ab % q * ga % q * zt
              + ga * ga % q * zt % q * (ht + 2 * r * at)) % q
    g, h = G1.mul(c.g1, td.g1_k), G2.mul(c.g2, td.g2_k)
    return G1.mul(g, a_dlog), G2.mul(h, a_dlog), G1.mul(g, c_dlog)


def verify(c: CurveParams, pk: Gm17ProvingKey, public_inputs: List[int], proof) -> bool:
    """ark-gm17 verify_proof:  e(A + G^alpha, B + H^beta) = e(G^alpha, H^beta) e(psi, H^gamma) e(C, H)  and  e(A, H^gamma) = e(G^gamma, B)."""
    G1, G2 = g1_group(c), g2_group(c)
    A, B, C = proof
    assert len(public_inputs) + 1 == len(pk.query)
    psi = pk.query[0]
    for x, P in zip(public_inputs, pk.query[1:]):
        psi = G1.add(psi, G1.mul(P, x))
    t1 = pairing_product_is_one(c, [(G1.add(A, pk.g_alpha_g1), G2.add(B, pk.h_beta_g2)), (G1.neg(pk.g_alpha_g1), pk.h_beta_g2),
                                    (G1.neg(psi), pk.h_gamma_g2), (G1.neg(C), pk.h_g2)])
    t2 = pairing_product_is_one(c, [(A, pk.h_gamma_g2), (G1.neg(pk.g_gamma_g1), B)])
    return t1 and t2


# ---- ark `ProvingKey::serialize_unchecked` layout (struct field order; Vec = u64 LE count + elements; uncompressed points) ----
def pk_serialize(c: CurveParams, pk: Gm17ProvingKey) -> bytes:
    def vec(items, ser):
        return len(items).to_bytes(8, "little") + b"".join(ser(c, p) for p in items)
    return (ser_g2(c, pk.h_g2) + ser_g1(c, pk.g_alpha_g1) + ser_g2(c, pk.h_beta_g2) + ser_g1(c, pk.g_gamma_g1) + ser_g2(c, pk.h_gamma_g2)
            + vec(pk.query, ser_g1) + vec(pk.a_query, ser_g1) + vec(pk.b_query, ser_g2) + vec(pk.c_query_1, ser_g1)
            + vec(pk.c_query_2, ser_g1) + ser_g1(c, pk.g_gamma_z) + ser_g2(c, pk.h_gamma_z) + ser_g1(c, pk.g_ab_gamma_z)
            + ser_g1(c, pk.g_gamma2_z2) + vec(pk.g_gamma2_z_t, ser_g1))


def pk_deserialize(c: CurveParams, data: bytes) -> Gm17ProvingKey:
    off = 0

    def one(de):
        nonlocal off
        p, off = de(c, data, off)
        return p

    def vec(de):
        nonlocal off
        n = int.from_bytes(data[off:off + 8], "little")
        off += 8
        return [one(de) for _ in range(n)]
    pk = Gm17ProvingKey(one(de_g2), one(de_g1), one(de_g2), one(de_g1), one(de_g2), vec(de_g1), vec(de_g1), vec(de_g2), vec(de_g1),
                        vec(de_g1), one(de_g1), one(de_g2), one(de_g1), one(de_g1), vec(de_g1))
    assert off == len(data)
    return pk
