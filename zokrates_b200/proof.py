"""Proof / key JSON types of the drop-in boundary.

Mirrors /root/reference/zokrates_proof_systems/src/lib.rs:19-96 (`SetupKeypair`, `Proof`, `G1Affine`,
`G2AffineFq2`), src/scheme/groth16.rs:8-35 (`G16`, `ProofPoints`, `VerificationKey`) and
src/tagged.rs:7-37 (`TaggedProof`, `TaggedVerificationKey`).  Point coordinates are "0x" + big-endian
hex, zero padded to the base-field byte length, exactly what `parse_g1`/`parse_g2`/`parse_fr`
(zokrates_ark/src/lib.rs:150-226) produce from ark's little-endian bytes.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import List, Tuple

from .curves import Curve, curve as _curve

SCHEME_NAME = "g16"     # `impl Scheme<T> for G16 { const NAME = "g16" }`  scheme/groth16.rs:27-28


def _hex(le_bytes: bytes) -> str:
    return "0x" + le_bytes[::-1].hex()


def _unhex(s: str, nbytes: int) -> bytes:
    b = bytes.fromhex(s[2:] if s.startswith("0x") else s)
    if len(b) > nbytes:
        raise ValueError("hex value too long")
    return b.rjust(nbytes, b"\0")[::-1]


@dataclass
class G1Affine:
    x: str
    y: str

    def to_json(self):
        return [self.x, self.y]


@dataclass
class G2Affine:            # G2AffineFq2((x.c0, x.c1), (y.c0, y.c1))
    x: Tuple[str, str]
    y: Tuple[str, str]

    def to_json(self):
        return [list(self.x), list(self.y)]


@dataclass
class ProofPoints:
    a: G1Affine
    b: G2Affine
    c: G1Affine

    def to_json(self):
        return {"a": self.a.to_json(), "b": self.b.to_json(), "c": self.c.to_json()}


@dataclass
class Proof:
    proof: ProofPoints
    inputs: List[str]
    curve: str = "bn128"
    scheme: str = SCHEME_NAME      # "g16" (scheme/groth16.rs) or "gm17" (scheme/gm17.rs:31): same a / b / c point layout

    @classmethod
    def from_raw(cls, c: Curve, raw: bytes, inputs: List[int], scheme: str = SCHEME_NAME) -> "Proof":
        """raw = A.x|A.y|B.x.c0|B.x.c1|B.y.c0|B.y.c1|C.x|C.y canonical LE (zkb_groth16_prove output)."""
        n = c.fq_bytes
        if len(raw) != 8 * n:
            raise ValueError("bad proof length")
        f = [_hex(raw[i * n:(i + 1) * n]) for i in range(8)]
        pts = ProofPoints(G1Affine(f[0], f[1]), G2Affine((f[2], f[3]), (f[4], f[5])), G1Affine(f[6], f[7]))
        return cls(pts, [_hex(int(v).to_bytes(c.fr_bytes, "little")) for v in inputs], c.name, scheme)

    def to_raw(self) -> bytes:
        c = _curve(self.curve)
        n = c.fq_bytes
        p = self.proof
        return b"".join(_unhex(s, n) for s in (p.a.x, p.a.y, p.b.x[0], p.b.x[1], p.b.y[0], p.b.y[1], p.c.x, p.c.y))

    def input_values(self) -> List[int]:
        return [int(s, 16) for s in self.inputs]

    def to_tagged_json(self) -> str:
        """`serde_json::to_string_pretty(&TaggedProof::<T, S>::new(proof.proof, proof.inputs))`
        (zokrates_cli/src/ops/generate_proof.rs:188-194): keys scheme, curve, proof, inputs."""
        return json.dumps({"scheme": self.scheme, "curve": self.curve, "proof": self.proof.to_json(),
                           "inputs": self.inputs}, indent=2)

    @classmethod
    def from_json(cls, text: str) -> "Proof":
        d = json.loads(text)
        p = d["proof"]
        pts = ProofPoints(G1Affine(*p["a"]), G2Affine(tuple(p["b"][0]), tuple(p["b"][1])), G1Affine(*p["c"]))
        return cls(pts, list(d["inputs"]), d.get("curve", "bn128"), d.get("scheme", SCHEME_NAME))


@dataclass
class VerificationKey:
    alpha: G1Affine
    beta: G2Affine
    gamma: G2Affine
    delta: G2Affine
    gamma_abc: List[G1Affine]
    curve: str = "bn128"

    def to_tagged_json(self) -> str:
        """TaggedVerificationKey: scheme, curve, then the flattened vk (tagged.rs:7-13)."""
        return json.dumps({"scheme": SCHEME_NAME, "curve": self.curve, "alpha": self.alpha.to_json(),
                           "beta": self.beta.to_json(), "gamma": self.gamma.to_json(), "delta": self.delta.to_json(),
                           "gamma_abc": [g.to_json() for g in self.gamma_abc]}, indent=2)


    @classmethod
    def from_json(cls, text: str) -> "VerificationKey":
        d = json.loads(text)
        if d.get("scheme", SCHEME_NAME) != SCHEME_NAME:
            raise ValueError("verification key is not for scheme g16")
        g2 = lambda v: G2Affine(tuple(v[0]), tuple(v[1]))
        return cls(G1Affine(*d["alpha"]), g2(d["beta"]), g2(d["gamma"]), g2(d["delta"]), [G1Affine(*g) for g in d["gamma_abc"]],
                   d.get("curve", "bn128"))


@dataclass
class SetupKeypair:
    vk: VerificationKey
    pk: bytes


def vk_from_pk_bytes(c: Curve, pk: bytes) -> VerificationKey:
    """The `vk` prefix of ark's ProvingKey layout re-encoded as the reference's hex VK
    (zokrates_ark/src/groth16.rs:100-106)."""
    n = c.fq_bytes
    off = 0

    def g1():
        nonlocal off
        raw = bytearray(pk[off:off + 2 * n]); off += 2 * n
        raw[-1] &= 0x3F
        return G1Affine(_hex(bytes(raw[:n])), _hex(bytes(raw[n:])))

    def g2():
        nonlocal off
        raw = bytearray(pk[off:off + 4 * n]); off += 4 * n
        raw[-1] &= 0x3F
        f = [_hex(bytes(raw[i * n:(i + 1) * n])) for i in range(4)]
        return G2Affine((f[0], f[1]), (f[2], f[3]))

    alpha, beta, gamma, delta = g1(), g2(), g2(), g2()
    cnt = int.from_bytes(pk[off:off + 8], "little"); off += 8
    abc = [g1() for _ in range(cnt)]
    return VerificationKey(alpha, beta, gamma, delta, abc, c.name)


@dataclass
class Gm17VerificationKey:
    """`VerificationKey<G1, G2>` of scheme/gm17.rs:19-26: h, g_alpha, h_beta, g_gamma, h_gamma, query."""
    h: G2Affine
    g_alpha: G1Affine
    h_beta: G2Affine
    g_gamma: G1Affine
    h_gamma: G2Affine
    query: List[G1Affine]
    curve: str = "bn128"

    def to_tagged_json(self) -> str:
        return json.dumps({"scheme": "gm17", "curve": self.curve, "h": self.h.to_json(), "g_alpha": self.g_alpha.to_json(),
                           "h_beta": self.h_beta.to_json(), "g_gamma": self.g_gamma.to_json(), "h_gamma": self.h_gamma.to_json(),
                           "query": [g.to_json() for g in self.query]}, indent=2)


def gm17_vk_from_pk_bytes(c: Curve, pk: bytes) -> Gm17VerificationKey:
    """The `vk` prefix of ark-gm17's ProvingKey layout re-encoded as the reference's hex VK (zokrates_ark/src/gm17.rs:29-36)."""
    n = c.fq_bytes
    off = 0

    def g1():
        nonlocal off
        raw = bytearray(pk[off:off + 2 * n]); off += 2 * n
        raw[-1] &= 0x3F
        return G1Affine(_hex(bytes(raw[:n])), _hex(bytes(raw[n:])))

    def g2():
        nonlocal off
        raw = bytearray(pk[off:off + 4 * n]); off += 4 * n
        raw[-1] &= 0x3F
        f = [_hex(bytes(raw[i * n:(i + 1) * n])) for i in range(4)]
        return G2Affine((f[0], f[1]), (f[2], f[3]))

    h, g_alpha, h_beta, g_gamma, h_gamma = g2(), g1(), g2(), g1(), g2()
    cnt = int.from_bytes(pk[off:off + 8], "little"); off += 8
    return Gm17VerificationKey(h, g_alpha, h_beta, g_gamma, h_gamma, [g1() for _ in range(cnt)], c.name)
