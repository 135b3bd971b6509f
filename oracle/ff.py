"""Big-int field / curve / pairing restatement (oracle — test infrastructure only).

Follows the arithmetic that `zokrates_field::FieldPrime` delegates to ark-ff 0.3.0
(`/root/reference/zokrates_field/src/lib.rs:407-503`: add/sub/mul/div/pow are thin wrappers
over `ark_ff::Fp256`), and the curve crates ark-bn254 / ark-bls12-381 0.3.0 pinned at
`/root/reference/Cargo.lock:91,102` (not vendored: constants restated in SURVEY.md App. C and
re-verified numerically by `tests/test_oracle_pins.py`).

Pure Python: small cases only.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class CurveParams:
    name: str            # zokrates Field::name()  (zokrates_field/src/bn128.rs:1-13)
    r: int               # scalar field modulus
    p: int               # base field modulus
    fr_bytes: int
    fq_bytes: int
    two_adicity: int
    fr_generator: int    # Fr::multiplicative_generator()
    b1: int              # G1: y^2 = x^3 + b1
    xi: tuple            # Fq2 non-residue used for the sextic twist
    twist: str           # 'D' (b2 = b1/xi) or 'M' (b2 = b1*xi)
    g1: tuple
    g2: tuple            # ((x.c0,x.c1),(y.c0,y.c1))
    ate_loop: int
    repr_shave_bits: int  # ark FpParameters::REPR_SHAVE_BITS of Fr
    bn_like: bool

    @property
    def b2(self):
        f = Fq2Ops(self.p)
        if self.twist == 'D':
            return f.mul((self.b1, 0), f.inv(self.xi))
        return f.mul((self.b1, 0), self.xi)

    @property
    def two_adic_root(self):
        return pow(self.fr_generator, (self.r - 1) >> self.two_adicity, self.r)


BN254 = CurveParams(
    name="bn128",
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    p=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    fr_bytes=32, fq_bytes=32, two_adicity=28, fr_generator=5,
    b1=3, xi=(9, 1), twist='D',
    g1=(1, 2),
    g2=((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    ate_loop=29793968203157093288, repr_shave_bits=2, bn_like=True,
)

BLS12_381 = CurveParams(
    name="bls12_381",
    r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    p=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    fr_bytes=32, fq_bytes=48, two_adicity=32, fr_generator=7,
    b1=4, xi=(1, 1), twist='M',
    g1=(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    g2=((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
         0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
        (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
         0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    ate_loop=15132376222941642752, repr_shave_bits=1, bn_like=False,
)

CURVES = {"bn128": BN254, "bls12_381": BLS12_381}
CURVE_IDS = {"bn128": 0, "bls12_381": 1}


def inv_mod(a: int, m: int) -> int:
    return pow(a, -1, m)


# --------------------------------------------------------------------------- Fq2
class Fq2Ops:
    """Fq2 = Fq[u]/(u^2+1) for both curves."""

    def __init__(self, p):
        self.p = p

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def neg(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a):
        return self.mul(a, a)

    def inv(self, a):
        p = self.p
        d = inv_mod((a[0] * a[0] + a[1] * a[1]) % p, p)
        return (a[0] * d % p, (-a[1]) * d % p)

    def is_zero(self, a):
        return a[0] % self.p == 0 and a[1] % self.p == 0

    zero = (0, 0)
    one = (1, 0)

    def from_int(self, k):
        return (k % self.p, 0)


class FqOps:
    def __init__(self, p):
        self.p = p

    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def neg(self, a):
        return (-a) % self.p

    def mul(self, a, b):
        return a * b % self.p

    def sqr(self, a):
        return a * a % self.p

    def inv(self, a):
        return inv_mod(a, self.p)

    def is_zero(self, a):
        return a % self.p == 0

    zero = 0
    one = 1

    def from_int(self, k):
        return k % self.p


# --------------------------------------------------------------------------- short Weierstrass, a = 0
class Group:
    """Affine / Jacobian arithmetic on y^2 = x^3 + b over a field given by `F` ops.

    Affine points are (x, y) tuples, the point at infinity is None (ark: `infinity` flag).
    """

    def __init__(self, F, b, order):
        self.F = F
        self.b = b
        self.order = order

    def is_on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sub(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b)) == F.zero

    def neg(self, P):
        if P is None:
            return None
        return (P[0], self.F.neg(P[1]))

    # Jacobian (X, Y, Z), infinity = Z == 0
    def to_jac(self, P):
        F = self.F
        if P is None:
            return (F.one, F.one, F.zero)
        return (P[0], P[1], F.one)

    def to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jdouble(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return J
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(t, t)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C)
        C8 = F.add(C8, C8)
        C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def jadd(self, J1, J2):
        F = self.F
        if F.is_zero(J1[2]):
            return J2
        if F.is_zero(J2[2]):
            return J1
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jdouble(J1)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        R = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(R), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def add(self, P, Q):
        return self.to_affine(self.jadd(self.to_jac(P), self.to_jac(Q)))

    def jmul(self, J, k):
        k %= self.order
        R = (self.F.one, self.F.one, self.F.zero)
        for bit in bin(k)[2:] if k else "":
            R = self.jdouble(R)
            if bit == "1":
                R = self.jadd(R, J)
        return R

    def mul(self, P, k):
        return self.to_affine(self.jmul(self.to_jac(P), k))

    def msm_naive(self, points, scalars):
        """Sum s_i * P_i by double-and-add — independent of any Pippenger code."""
        acc = (self.F.one, self.F.one, self.F.zero)
        for P, s in zip(points, scalars):
            if P is None or s % self.order == 0:
                continue
            acc = self.jadd(acc, self.jmul(self.to_jac(P), s))
        return self.to_affine(acc)


def g1_group(c: CurveParams) -> Group:
    return Group(FqOps(c.p), c.b1, c.r)


def g2_group(c: CurveParams) -> Group:
    return Group(Fq2Ops(c.p), c.b2, c.r)


# --------------------------------------------------------------------------- Fq12 as Fq[w]/(w^12 - 2*xi0*w^6 + |xi|^2)
class Fq12Poly:
    """Degree-12 extension as a polynomial ring: w^6 = xi = xi0 + xi1*u  (u^2 = -1), so
    (w^6 - xi0)^2 = -xi1^2  =>  w^12 = 2*xi0*w^6 - (xi0^2 + xi1^2).
    Only used by the pairing check (reference: ark `verify_proof`, reached from
    /root/reference/zokrates_ark/src/groth16.rs:85) — any non-degenerate bilinear map decides the
    same equation, so the exact Miller-loop normalisation of ark is not restated.
    """

    def __init__(self, c: CurveParams):
        self.p = c.p
        self.xi0, self.xi1 = c.xi
        assert self.xi1 == 1
        self.c6 = (2 * self.xi0) % c.p
        self.c0 = (-(self.xi0 * self.xi0 + 1)) % c.p

    def one(self):
        return [1] + [0] * 11

    def from_fq2(self, a):
        # a0 + a1*u, u = w^6 - xi0
        v = [0] * 12
        v[0] = (a[0] - self.xi0 * a[1]) % self.p
        v[6] = a[1] % self.p
        return v

    def from_fq(self, a):
        v = [0] * 12
        v[0] = a % self.p
        return v

    def add(self, a, b):
        p = self.p
        return [(x + y) % p for x, y in zip(a, b)]

    def sub(self, a, b):
        p = self.p
        return [(x - y) % p for x, y in zip(a, b)]

    def mul(self, a, b):
        p = self.p
        t = [0] * 23
        for i, x in enumerate(a):
            if x:
                for j, y in enumerate(b):
                    t[i + j] += x * y
        for k in range(22, 11, -1):
            v = t[k]
            if v:
                t[k - 6] += v * self.c6
                t[k - 12] += v * self.c0
        return [x % p for x in t[:12]]

    def scalar(self, a, k):
        return [x * k % self.p for x in a]

    def pow(self, a, e):
        r = self.one()
        for bit in bin(e)[2:]:
            r = self.mul(r, r)
            if bit == "1":
                r = self.mul(r, a)
        return r

    def inv(self, a):
        # extended Euclid over Fq[w]
        p = self.p
        mod = [self.c0 * -1 % p] + [0] * 5 + [(-self.c6) % p] + [0] * 5 + [1]

        def deg(v):
            d = len(v) - 1
            while d >= 0 and v[d] % p == 0:
                d -= 1
            return d

        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = list(a) + [0], mod
        while deg(low) > 0:
            dl, dh = deg(low), deg(high)
            # r = high // low
            temp = list(high)
            q = [0] * 13
            il = inv_mod(low[dl], p)
            for i in range(dh - dl, -1, -1):
                q[i] = temp[dl + i] * il % p
                if q[i]:
                    for j in range(dl + 1):
                        temp[i + j] = (temp[i + j] - q[i] * low[j]) % p
            nm = list(hm)
            new = temp
            for i in range(13):
                if lm[i]:
                    for j in range(13 - i):
                        if q[j]:
                            nm[i + j] = (nm[i + j] - lm[i] * q[j]) % p
            lm, low, hm, high = nm, new, lm, low
        il = inv_mod(low[0], p)
        return [x * il % p for x in lm[:12]]


def _miller_and_points(c: CurveParams, P, Q):
    """Generic-line Miller loop in Fq12 over the untwisted image of Q."""
    K = Fq12Poly(c)
    w = [0, 1] + [0] * 10
    w2 = K.mul(w, w)
    w3 = K.mul(w2, w)
    qx, qy = K.from_fq2(Q[0]), K.from_fq2(Q[1])
    if c.twist == 'D':
        Qx, Qy = K.mul(qx, w2), K.mul(qy, w3)
    else:
        Qx, Qy = K.mul(qx, K.inv(w2)), K.mul(qy, K.inv(w3))
    Px, Py = K.from_fq(P[0]), K.from_fq(P[1])

    def line(P1, P2, T):
        x1, y1 = P1
        x2, y2 = P2
        xt, yt = T
        if x1 != x2:
            m = K.mul(K.sub(y2, y1), K.inv(K.sub(x2, x1)))
            return K.sub(K.mul(m, K.sub(xt, x1)), K.sub(yt, y1))
        if y1 == y2:
            m = K.mul(K.scalar(K.mul(x1, x1), 3), K.inv(K.scalar(y1, 2)))
            return K.sub(K.mul(m, K.sub(xt, x1)), K.sub(yt, y1))
        return K.sub(xt, x1)

    def padd(P1, P2):
        x1, y1 = P1
        x2, y2 = P2
        if x1 == x2 and y1 == y2:
            m = K.mul(K.scalar(K.mul(x1, x1), 3), K.inv(K.scalar(y1, 2)))
        else:
            m = K.mul(K.sub(y2, y1), K.inv(K.sub(x2, x1)))
        x3 = K.sub(K.sub(K.mul(m, m), x1), x2)
        y3 = K.sub(K.mul(m, K.sub(x1, x3)), y1)
        return (x3, y3)

    Qp = (Qx, Qy)
    R = Qp
    f = K.one()
    Pt = (Px, Py)
    for bit in bin(c.ate_loop)[3:]:
        f = K.mul(K.mul(f, f), line(R, R, Pt))
        R = padd(R, R)
        if bit == "1":
            f = K.mul(f, line(R, Qp, Pt))
            R = padd(R, Qp)
    if c.bn_like:
        Q1 = (K.pow(Qx, c.p), K.pow(Qy, c.p))
        nQ2 = (K.pow(Q1[0], c.p), K.sub([0] * 12, K.pow(Q1[1], c.p)))
        f = K.mul(f, line(R, Q1, Pt))
        R = padd(R, Q1)
        f = K.mul(f, line(R, nQ2, Pt))
    return K, f


def miller_loop(c: CurveParams, P, Q):
    """Miller function value (before final exponentiation) for P in G1, Q in G2 (affine, not None)."""
    if P is None or Q is None:
        return Fq12Poly(c).one()
    return _miller_and_points(c, P, Q)[1]


def final_exp(c: CurveParams, f):
    K = Fq12Poly(c)
    return K.pow(f, (c.p ** 12 - 1) // c.r)


def pairing_product_is_one(c: CurveParams, pairs) -> bool:
    """prod e(P_i, Q_i) == 1 — the form ark's verifier checks (one final exponentiation)."""
    K = Fq12Poly(c)
    f = K.one()
    for P, Q in pairs:
        f = K.mul(f, miller_loop(c, P, Q))
    return final_exp(c, f) == K.one()
