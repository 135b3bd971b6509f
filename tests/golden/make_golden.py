#!/usr/bin/env python3
"""Regenerates tests/golden/*.json from the CPU oracle (oracle/ff.py, oracle/ark.py — python big-int restatement of the
reference's algorithm; test infrastructure only).

The reference cannot run here (Rust workspace, un-vendored arkworks crates, no cargo), so these fixtures are
ORACLE-generated and pinned to the reference only through what the reference's own tests pin (field / format KATs in
tests/test_oracle_pins.py) plus the algebraic checks recorded below (every proof here satisfies the pairing equation and equals
the trapdoor prediction).  Proof VALUES against zokrates_ark remain "parity unpinned" (DESIGN.md §2).
The fixtures freeze today's agreed answers: the GPU tier, the host-emulated engine and the oracle itself are all compared with
them, so a regression in any one of the three shows up without the other two moving along.

    python tests/golden/make_golden.py            # rewrites the JSON files next to this script
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ark, ir as oir                      # noqa: E402
from oracle.ff import BLS12_381, BN254, g1_group, g2_group  # noqa: E402

CURVES = {"bn128": BN254, "bls12_381": BLS12_381}


def ntt_vectors():
    out = []
    for name, c in CURVES.items():
        for log_n in (3, 5):
            rnd = random.Random(0x60D + log_n)
            x = [rnd.randrange(c.r) for _ in range(1 << log_n)]
            d = ark.Domain(c, 1 << log_n)
            out.append({"curve": name, "log_n": log_n, "x": [str(v) for v in x], "fft": [str(v) for v in d.fft(x)],
                        "ifft": [str(v) for v in d.ifft(x)], "coset_fft": [str(v) for v in d.coset_fft(x)],
                        "coset_ifft": [str(v) for v in d.coset_ifft(x)]})
    return out


def msm_vectors():
    out = []
    for name, c in CURVES.items():
        rnd = random.Random(0x35)
        G1, G2 = g1_group(c), g2_group(c)
        p1 = [G1.mul(c.g1, rnd.randrange(1, c.r)) for _ in range(6)]
        p2 = [G2.mul(c.g2, rnd.randrange(1, c.r)) for _ in range(5)]
        p1[2] = None                                   # point at infinity
        p1[4] = p1[3]                                  # the same point twice
        s1 = [0, 1, rnd.randrange(c.r), c.r - 1, c.r - 1, rnd.randrange(1 << 20)]
        s2 = [rnd.randrange(c.r), 1, 0, c.r - 1, 2]
        out.append({"curve": name,
                    "g1_points": b"".join(ark.ser_g1(c, p) for p in p1).hex(), "g1_scalars": [str(v) for v in s1],
                    "g1_result": ark.ser_g1(c, G1.msm_naive(p1, s1)).hex(),
                    "g2_points": b"".join(ark.ser_g2(c, p) for p in p2).hex(), "g2_scalars": [str(v) for v in s2],
                    "g2_result": ark.ser_g2(c, G2.msm_naive(p2, s2)).hex()})
    return out


def rng_vectors():
    out = []
    for entropy in ("", "smoke", "some entropy"):
        for name, c in CURVES.items():
            rng = ark.rng_from_entropy(entropy)
            out.append({"entropy": entropy, "curve": name, "fr_rand": [str(ark.fr_rand(c, rng)) for _ in range(3)]})
    return out


def proof_vectors():
    """BASELINE config 1 (`def main(private field a, field b) { assert(a * a == b); }`, 337 113569) and a program with a
    return value, both curves: setup with an explicit trapdoor, proof with r, s from the entropy string."""
    progs = {
        "factorize": (oir.Prog([(1, True), (2, False)], 0, [oir.Constraint([(1, 1)], [(1, 1)], [(2, 1)])]), [337, 113569]),
        "with_output": (oir.Prog([(1, True), (2, False)], 1, [oir.Constraint([(1, 1), (0, 5)], [(2, 3)], [(3, 1)]),
                                                              oir.Constraint([(3, 1)], [(1, 1)], [(-1, 1)])]), [7, 11]),
    }
    td = [11, 22, 33, 44, 55555, 3, 7]
    out = []
    for name, c in CURVES.items():
        for pname, (prog, inputs) in progs.items():
            w = oir.execute(c, prog, inputs)
            r1cs, z = ark.synthesize(prog, w)
            pk = ark.setup(c, r1cs, ark.Trapdoor(*td))
            pk_bytes = ark.pk_serialize(c, pk)
            rng = ark.rng_from_entropy("golden")
            r, s = ark.fr_rand(c, rng), ark.fr_rand(c, rng)
            proof = ark.prove(c, pk, r1cs, z, r, s)
            expected = ark.trapdoor_expected_proof(c, r1cs, ark.Trapdoor(*td), z, r, s)
            pub = prog.public_inputs_values(w)
            assert ark.tagged_proof_json(c, proof, pub) == ark.tagged_proof_json(c, expected, pub)
            assert ark.verify(c, pk, pub, proof)
            out.append({"curve": name, "program": pname, "inputs": [str(v) for v in inputs], "trapdoor": td, "entropy": "golden",
                        "arguments": [[v, priv] for v, priv in prog.arguments], "return_count": prog.return_count,
                        "constraints": [[c_.left, c_.right, c_.lin] for c_ in prog.statements],
                        "assignment": [str(v) for v in z], "r": str(r), "s": str(s),
                        "pk_sha256": hashlib.sha256(pk_bytes).hexdigest(), "pk_len": len(pk_bytes),
                        "proof_json": ark.tagged_proof_json(c, proof, pub), "pairing_check": True, "equals_trapdoor_prediction": True})
    return out


def main():
    for fname, fn in (("ntt.json", ntt_vectors), ("msm.json", msm_vectors), ("rng.json", rng_vectors), ("proofs.json", proof_vectors)):
        with open(os.path.join(HERE, fname), "w") as f:
            json.dump(fn(), f, indent=1)
            f.write("\n")
        print("wrote", fname)


if __name__ == "__main__":
    main()
