"""arkworks-0.3.0 / zokrates_ark Groth16 restatement in big-int Python — oracle, test infrastructure only.

What it restates:
  * get_rng_from_entropy                /root/reference/zokrates_proof_systems/src/rng.rs:5-20
  * rand 0.8.5 StdRng (= ChaCha12), ark-ff 0.3.0 `Fp::rand`     (external, Cargo.lock:2447,2458,161)
  * Computation::generate_constraints   /root/reference/zokrates_ark/src/lib.rs:76-130 (variable order)
  * ark-groth16 0.3.0 create_random_proof / witness_map / generate_parameters / verify_proof
    (external, Cargo.lock:221; call sites zokrates_ark/src/groth16.rs:41,44,85,95)
  * ark-serialize 0.3.0 ProvingKey::serialize_unchecked layout  (call site groth16.rs:97-98)
  * parse_g1 / parse_g2 / parse_fr      /root/reference/zokrates_ark/src/lib.rs:150-226
  * TaggedProof JSON                    /root/reference/zokrates_proof_systems/src/tagged.rs:14-37

Proof values: "parity unpinned" (no golden proof exists in the reference) — validated by the
pairing equation and by the trapdoor check below, both independent of the MSM/NTT code.
"""
from __future__ import annotations

import hashlib
import json
import struct
from dataclasses import dataclass
from typing import List, Optional

from .ff import CurveParams, g1_group, g2_group, inv_mod, pairing_product_is_one
from .ir import Constraint, Prog

MASK32 = 0xFFFFFFFF


# ----------------------------------------------------------------------------- RNG
class ChaCha12Rng:
    """rand_chacha 0.3.1 ChaCha12Rng::from_seed: key = seed, 64-bit block counter 0, stream 0;
    words are consumed sequentially, next_u64 = (lo word, hi word)."""

    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = list(struct.unpack("<8I", seed))
        self.counter = 0
        self.buf: List[int] = []

    @staticmethod
    def _qr(s, a, b, c, d):
        s[a] = (s[a] + s[b]) & MASK32; s[d] ^= s[a]; s[d] = ((s[d] << 16) | (s[d] >> 16)) & MASK32
        s[c] = (s[c] + s[d]) & MASK32; s[b] ^= s[c]; s[b] = ((s[b] << 12) | (s[b] >> 20)) & MASK32
        s[a] = (s[a] + s[b]) & MASK32; s[d] ^= s[a]; s[d] = ((s[d] << 8) | (s[d] >> 24)) & MASK32
        s[c] = (s[c] + s[d]) & MASK32; s[b] ^= s[c]; s[b] = ((s[b] << 7) | (s[b] >> 25)) & MASK32

    def _block(self):
        init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + self.key + [
            self.counter & MASK32, (self.counter >> 32) & MASK32, 0, 0]
        s = list(init)
        for _ in range(6):
            self._qr(s, 0, 4, 8, 12); self._qr(s, 1, 5, 9, 13); self._qr(s, 2, 6, 10, 14); self._qr(s, 3, 7, 11, 15)
            self._qr(s, 0, 5, 10, 15); self._qr(s, 1, 6, 11, 12); self._qr(s, 2, 7, 8, 13); self._qr(s, 3, 4, 9, 14)
        self.counter += 1
        self.buf.extend((x + y) & MASK32 for x, y in zip(s, init))

    def next_u32(self) -> int:
        if not self.buf:
            self._block()
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


def rng_from_entropy(entropy: str) -> ChaCha12Rng:
    """rng.rs:5-20 — first 32 bytes of Blake2b-512(entropy)."""
    return ChaCha12Rng(hashlib.blake2b(entropy.encode(), digest_size=64).digest()[:32])


def fr_rand(c: CurveParams, rng: ChaCha12Rng) -> int:
    """ark-ff 0.3.0 `impl Distribution<Fp256<P>> for Standard`: 4 limbs from next_u64, mask the top
    REPR_SHAVE_BITS, accept if < modulus; the accepted integer IS the Montgomery representation
    (SURVEY.md App. B.5), so the field value is limbs * R^-1."""
    while True:
        limbs = [rng.next_u64() for _ in range(4)]
        limbs[3] &= (0xFFFFFFFFFFFFFFFF >> c.repr_shave_bits)
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < c.r:
            return v * inv_mod(1 << 256, c.r) % c.r


# ----------------------------------------------------------------------------- R1CS in ark order
@dataclass
class R1CS:
    """Matrices in ark-relations column order: 0 = one, 1..l = instance, then witness."""
    num_instance: int          # incl. one
    num_witness: int
    a: list                    # per constraint: [(col, coeff)]
    b: list
    c: list

    @property
    def num_constraints(self):
        return len(self.a)

    @property
    def num_variables(self):
        return self.num_instance + self.num_witness


def synthesize(prog: Prog, witness: Optional[dict] = None):
    """Computation::generate_constraints  zokrates_ark/src/lib.rs:80-130.

    Returns (R1CS, z) with z the full assignment [1, instance.., witness..] (None without witness).
    Instance / witness variables are numbered separately in allocation order; public args and
    outputs (`id < 0`, lib.rs:52) are instance variables.
    """
    symbols = {0: ("i", 0)}
    inst = [0]           # variable ids, allocation order
    wit = []
    for v, private in prog.arguments:             # lib.rs:94-113
        if private:
            symbols[v] = ("w", len(wit)); wit.append(v)
        else:
            symbols[v] = ("i", len(inst)); inst.append(v)

    rows = []

    def comb(l):                                   # ark_combination lib.rs:41-74
        out = []
        for v, coeff in l:
            if v not in symbols:
                if v < 0:
                    symbols[v] = ("i", len(inst)); inst.append(v)
                else:
                    symbols[v] = ("w", len(wit)); wit.append(v)
            out.append((symbols[v], coeff))
        return out

    for s in prog.statements:                      # lib.rs:115-123 (directives / logs skipped)
        if isinstance(s, Constraint):
            rows.append((comb(s.left), comb(s.right), comb(s.lin)))

    ni = len(inst)

    def col(sym):
        return sym[1] if sym[0] == "i" else ni + sym[1]

    a = [[(col(s), k) for s, k in r[0]] for r in rows]
    b = [[(col(s), k) for s, k in r[1]] for r in rows]
    cc = [[(col(s), k) for s, k in r[2]] for r in rows]
    z = None
    if witness is not None:
        z = [witness[v] for v in inst] + [witness[v] for v in wit]
    return R1CS(ni, len(wit), a, b, cc), z


# ----------------------------------------------------------------------------- domain / FFT
class Domain:
    """ark-poly 0.3.0 Radix2EvaluationDomain::new(num_coeffs)."""

    def __init__(self, c: CurveParams, num_coeffs: int):
        self.c = c
        n = 1
        while n < num_coeffs:
            n <<= 1
        self.n = n
        self.log_n = n.bit_length() - 1
        assert self.log_n <= c.two_adicity
        self.omega = pow(c.two_adic_root, 1 << (c.two_adicity - self.log_n), c.r)
        self.g = c.fr_generator

    def fft(self, a, inverse=False):
        r = self.c.r
        n = self.n
        a = list(a) + [0] * (n - len(a))
        w = inv_mod(self.omega, r) if inverse else self.omega
        # bit reversal + iterative DIT
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            wl = pow(w, n // length, r)
            for s in range(0, n, length):
                wk = 1
                for k in range(length // 2):
                    u = a[s + k]
                    v = a[s + k + length // 2] * wk % r
                    a[s + k] = (u + v) % r
                    a[s + k + length // 2] = (u - v) % r
                    wk = wk * wl % r
            length <<= 1
        if inverse:
            ninv = inv_mod(n, r)
            a = [x * ninv % r for x in a]
        return a

    def ifft(self, a):
        return self.fft(a, True)

    def coset_fft(self, a):
        r = self.c.r
        gp, out = 1, []
        for x in a:
            out.append(x * gp % r)
            gp = gp * self.g % r
        return self.fft(out)

    def coset_ifft(self, a):
        r = self.c.r
        a = self.ifft(a)
        gi = inv_mod(self.g, r)
        gp, out = 1, []
        for x in a:
            out.append(x * gp % r)
            gp = gp * gi % r
        return out

    def lagrange_at(self, t):
        """evaluate_all_lagrange_coefficients(t): L_j(t) = Z(t)/n * w^j/(t - w^j)."""
        r = self.c.r
        zt = (pow(t, self.n, r) - 1) % r
        out, wj = [], 1
        ninv = inv_mod(self.n, r)
        for _ in range(self.n):
            out.append(zt * ninv % r * wj % r * inv_mod((t - wj) % r, r) % r)
            wj = wj * self.omega % r
        return out


def eval_row(c, row, z):
    acc = 0
    for col, k in row:
        acc += k * z[col]
    return acc % c.r


def witness_map(c: CurveParams, r1cs: R1CS, z: list) -> list:
    """ark-groth16 0.3.0 LibsnarkReduction::witness_map."""
    N, ni = r1cs.num_constraints, r1cs.num_instance
    d = Domain(c, N + ni)
    n, r = d.n, c.r
    a = [eval_row(c, row, z) for row in r1cs.a] + [0] * (n - N)
    b = [eval_row(c, row, z) for row in r1cs.b] + [0] * (n - N)
    for j in range(ni):
        a[N + j] = z[j]
    a = d.coset_fft(d.ifft(a))
    b = d.coset_fft(d.ifft(b))
    ab = [x * y % r for x, y in zip(a, b)]
    cc = [eval_row(c, row, z) for row in r1cs.c] + [0] * (n - N)
    cc = d.coset_fft(d.ifft(cc))
    zinv = inv_mod((pow(d.g, n, r) - 1) % r, r)
    ab = [(x - y) * zinv % r for x, y in zip(ab, cc)]
    return d.coset_ifft(ab)


# ----------------------------------------------------------------------------- keys
@dataclass
class ProvingKey:
    alpha_g1: tuple
    beta_g2: tuple
    gamma_g2: tuple
    delta_g2: tuple
    gamma_abc_g1: list
    beta_g1: tuple
    delta_g1: tuple
    a_query: list
    b_g1_query: list
    b_g2_query: list
    h_query: list
    l_query: list


@dataclass
class Trapdoor:
    alpha: int
    beta: int
    gamma: int
    delta: int
    tau: int
    g1_k: int = 1      # g1 generator = g1_k * standard generator
    g2_k: int = 1


def qap_at_tau(c: CurveParams, r1cs: R1CS, tau: int):
    """a_i(tau), b_i(tau), c_i(tau) per variable (ark generate_parameters, App. B.6)."""
    N, ni = r1cs.num_constraints, r1cs.num_instance
    d = Domain(c, N + ni)
    u = d.lagrange_at(tau)
    m, r = r1cs.num_variables, c.r
    a, b, cc = [0] * m, [0] * m, [0] * m
    for i in range(ni):
        a[i] = u[N + i]
    for j in range(N):
        for col, k in r1cs.a[j]:
            a[col] = (a[col] + u[j] * k) % r
        for col, k in r1cs.b[j]:
            b[col] = (b[col] + u[j] * k) % r
        for col, k in r1cs.c[j]:
            cc[col] = (cc[col] + u[j] * k) % r
    zt = (pow(tau, d.n, r) - 1) % r
    return d, a, b, cc, zt


def setup(c: CurveParams, r1cs: R1CS, td: Trapdoor) -> ProvingKey:
    G1, G2 = g1_group(c), g2_group(c)
    r = c.r
    d, a, b, cc, zt = qap_at_tau(c, r1cs, td.tau)
    ni = r1cs.num_instance
    g1 = G1.mul(c.g1, td.g1_k)
    g2 = G2.mul(c.g2, td.g2_k)
    ginv, dinv = inv_mod(td.gamma, r), inv_mod(td.delta, r)
    abc = [(td.beta * a[i] + td.alpha * b[i] + cc[i]) % r for i in range(r1cs.num_variables)]
    h_scalars, tp = [], 1
    for _ in range(d.n - 1):
        h_scalars.append(zt * dinv % r * tp % r)
        tp = tp * td.tau % r
    return ProvingKey(
        alpha_g1=G1.mul(g1, td.alpha), beta_g2=G2.mul(g2, td.beta), gamma_g2=G2.mul(g2, td.gamma),
        delta_g2=G2.mul(g2, td.delta),
        gamma_abc_g1=[G1.mul(g1, abc[i] * ginv % r) for i in range(ni)],
        beta_g1=G1.mul(g1, td.beta), delta_g1=G1.mul(g1, td.delta),
        a_query=[G1.mul(g1, x) for x in a],
        b_g1_query=[G1.mul(g1, x) for x in b],
        b_g2_query=[G2.mul(g2, x) for x in b],
        h_query=[G1.mul(g1, x) for x in h_scalars],
        l_query=[G1.mul(g1, abc[i] * dinv % r) for i in range(ni, r1cs.num_variables)],
    )


# ----------------------------------------------------------------------------- serialization (App. A.3)
def _fq_bytes(c, v):
    return int(v).to_bytes(c.fq_bytes, "little")


def ser_g1(c: CurveParams, P) -> bytes:
    if P is None:
        b = bytearray(2 * c.fq_bytes)
        b[-1] |= 0x40
        return bytes(b)
    return _fq_bytes(c, P[0]) + _fq_bytes(c, P[1])


def ser_g2(c: CurveParams, P) -> bytes:
    if P is None:
        b = bytearray(4 * c.fq_bytes)
        b[-1] |= 0x40
        return bytes(b)
    return b"".join(_fq_bytes(c, v) for v in (P[0][0], P[0][1], P[1][0], P[1][1]))


def de_g1(c, data, off):
    n = c.fq_bytes
    raw = bytearray(data[off:off + 2 * n])
    inf = bool(raw[-1] & 0x40)
    raw[-1] &= 0x3F
    if inf:
        return None, off + 2 * n
    return (int.from_bytes(raw[:n], "little"), int.from_bytes(raw[n:], "little")), off + 2 * n


def de_g2(c, data, off):
    n = c.fq_bytes
    raw = bytearray(data[off:off + 4 * n])
    inf = bool(raw[-1] & 0x40)
    raw[-1] &= 0x3F
    if inf:
        return None, off + 4 * n
    v = [int.from_bytes(raw[i * n:(i + 1) * n], "little") for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3])), off + 4 * n


def pk_serialize(c: CurveParams, pk: ProvingKey) -> bytes:
    """ProvingKey::serialize_unchecked: vk{alpha_g1,beta_g2,gamma_g2,delta_g2,gamma_abc_g1},
    beta_g1, delta_g1, a_query, b_g1_query, b_g2_query, h_query, l_query; Vec = u64 LE len + items."""
    out = [ser_g1(c, pk.alpha_g1), ser_g2(c, pk.beta_g2), ser_g2(c, pk.gamma_g2), ser_g2(c, pk.delta_g2)]

    def vec(items, ser):
        out.append(struct.pack("<Q", len(items)))
        out.extend(ser(c, p) for p in items)

    vec(pk.gamma_abc_g1, ser_g1)
    out.append(ser_g1(c, pk.beta_g1))
    out.append(ser_g1(c, pk.delta_g1))
    vec(pk.a_query, ser_g1)
    vec(pk.b_g1_query, ser_g1)
    vec(pk.b_g2_query, ser_g2)
    vec(pk.h_query, ser_g1)
    vec(pk.l_query, ser_g1)
    return b"".join(out)


def pk_deserialize(c: CurveParams, data: bytes) -> ProvingKey:
    off = 0
    alpha, off = de_g1(c, data, off)
    beta2, off = de_g2(c, data, off)
    gamma2, off = de_g2(c, data, off)
    delta2, off = de_g2(c, data, off)

    def vec(de, off):
        (n,) = struct.unpack_from("<Q", data, off)
        off += 8
        items = []
        for _ in range(n):
            p, off = de(c, data, off)
            items.append(p)
        return items, off

    abc, off = vec(de_g1, off)
    beta1, off = de_g1(c, data, off)
    delta1, off = de_g1(c, data, off)
    aq, off = vec(de_g1, off)
    b1q, off = vec(de_g1, off)
    b2q, off = vec(de_g2, off)
    hq, off = vec(de_g1, off)
    lq, off = vec(de_g1, off)
    assert off == len(data), "trailing bytes in proving key"
    return ProvingKey(alpha, beta2, gamma2, delta2, abc, beta1, delta1, aq, b1q, b2q, hq, lq)


# ----------------------------------------------------------------------------- prover (App. B.1)
def prove(c: CurveParams, pk: ProvingKey, r1cs: R1CS, z: list, r: int, s: int, msm1=None, msm2=None):
    """ark-groth16 0.3.0 create_proof_with_reduction.  Returns affine (A, B, C)."""
    G1, G2 = g1_group(c), g2_group(c)
    msm1 = msm1 or G1.msm_naive
    msm2 = msm2 or G2.msm_naive
    ni = r1cs.num_instance
    h = witness_map(c, r1cs, z)
    h_acc = msm1(pk.h_query, h)                       # zip => first n-1 coefficients
    aux = z[ni:]
    l_acc = msm1(pk.l_query, aux)
    assignment = z[1:]
    add1 = G1.add
    g_a = add1(add1(add1(G1.mul(pk.delta_g1, r), pk.a_query[0]), msm1(pk.a_query[1:], assignment)), pk.alpha_g1)
    if r % c.r != 0:
        g1_b = add1(add1(add1(G1.mul(pk.delta_g1, s), pk.b_g1_query[0]), msm1(pk.b_g1_query[1:], assignment)),
                    pk.beta_g1)
    else:
        g1_b = None
    g2_b = G2.add(G2.add(G2.add(G2.mul(pk.delta_g2, s), pk.b_g2_query[0]), msm2(pk.b_g2_query[1:], assignment)),
                  pk.beta_g2)
    g_c = G1.mul(g_a, s)
    g_c = add1(g_c, G1.mul(g1_b, r))
    g_c = add1(g_c, G1.neg(G1.mul(pk.delta_g1, r * s % c.r)))
    g_c = add1(add1(g_c, l_acc), h_acc)
    return g_a, g2_b, g_c


def trapdoor_expected_proof(c: CurveParams, r1cs: R1CS, td: Trapdoor, z: list, r: int, s: int):
    """Expected (A, B, C) from the trapdoor with Fr arithmetic + one scalar-mul each — no NTT, no MSM."""
    G1, G2 = g1_group(c), g2_group(c)
    q = c.r
    d, a, b, cc, zt = qap_at_tau(c, r1cs, td.tau)
    az = sum(x * y for x, y in zip(a, z)) % q
    bz = sum(x * y for x, y in zip(b, z)) % q
    cz = sum(x * y for x, y in zip(cc, z)) % q
    dinv = inv_mod(td.delta, q)
    a_dlog = (td.alpha + az + r * td.delta) % q
    b_dlog = (td.beta + bz + s * td.delta) % q
    ni = r1cs.num_instance
    l_dlog = sum((td.beta * a[i] + td.alpha * b[i] + cc[i]) * z[i] for i in range(ni, len(z))) % q * dinv % q
    h_dlog = (az * bz - cz) % q * dinv % q            # h(tau) * Z(tau) / delta
    c_dlog = (l_dlog + h_dlog + s * a_dlog + r * b_dlog - r * s % q * td.delta) % q
    g1 = G1.mul(c.g1, td.g1_k)
    g2 = G2.mul(c.g2, td.g2_k)
    return G1.mul(g1, a_dlog), G2.mul(g2, b_dlog), G1.mul(g1, c_dlog)


def verify(c: CurveParams, pk: ProvingKey, public_inputs: list, proof) -> bool:
    """ark verify_proof: e(A,B) == e(alpha,beta) * e(sum x_i gamma_abc_i, gamma) * e(C, delta)."""
    G1 = g1_group(c)
    A, B, C = proof
    assert len(public_inputs) + 1 == len(pk.gamma_abc_g1)
    acc = pk.gamma_abc_g1[0]
    for x, P in zip(public_inputs, pk.gamma_abc_g1[1:]):
        acc = G1.add(acc, G1.mul(P, x))
    return pairing_product_is_one(c, [(A, B), (G1.neg(pk.alpha_g1), pk.beta_g2), (G1.neg(acc), pk.gamma_g2),
                                      (G1.neg(C), pk.delta_g2)])


# ----------------------------------------------------------------------------- proof encoding (App. A.4)
def hex_fq(c, v):
    return "0x" + int(v).to_bytes(c.fq_bytes, "big").hex()


def hex_fr(c, v):
    return "0x" + int(v).to_bytes(c.fr_bytes, "big").hex()


def proof_points_json(c: CurveParams, proof) -> dict:
    A, B, C = proof
    return {"a": [hex_fq(c, A[0]), hex_fq(c, A[1])],
            "b": [[hex_fq(c, B[0][0]), hex_fq(c, B[0][1])], [hex_fq(c, B[1][0]), hex_fq(c, B[1][1])]],
            "c": [hex_fq(c, C[0]), hex_fq(c, C[1])]}


def tagged_proof_json(c: CurveParams, proof, inputs: list) -> str:
    """serde_json::to_string_pretty(TaggedProof) — generate_proof.rs:188-194."""
    return json.dumps({"scheme": "g16", "curve": c.name, "proof": proof_points_json(c, proof),
                       "inputs": [hex_fr(c, v) for v in inputs]}, indent=2)


def generate_proof(c: CurveParams, prog: Prog, witness: dict, pk_bytes: bytes, rng: ChaCha12Rng):
    """Backend::<T,G16>::generate_proof for Ark  — zokrates_ark/src/groth16.rs:21-53."""
    inputs = prog.public_inputs_values(witness)
    pk = pk_deserialize(c, pk_bytes)
    r = fr_rand(c, rng)
    s = fr_rand(c, rng)
    r1cs, z = synthesize(prog, witness)
    return prove(c, pk, r1cs, z, r, s), inputs
