#!/usr/bin/env python3
"""zkb-verify: `zokrates verify` (/root/reference/zokrates_cli/src/ops/verify.rs:17-59,175-196): checks `proof.json` against
`verification.key` with the product's host verifier (zokrates_b200/verify.py — verification is host work in the reference
too) and prints PASSED or FAILED.

    python tools/zkb_verify.py -j proof.json -v verification.key
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="zkb-verify", description="Verifies a given proof with the given verification key")
    ap.add_argument("-j", "--proof-path", default="proof.json", metavar="FILE", help="Path of the JSON proof path")
    ap.add_argument("-v", "--verification-key-path", default="verification.key", metavar="FILE", help="Path of the generated verification key file")
    ap.add_argument("-b", "--backend", default="b200", choices=["b200"], help="Backend to use")
    args = ap.parse_args(argv)

    from zokrates_b200 import backend, proof as pproof

    def slurp(path):
        try:
            with open(path) as f:
                return f.read()
        except OSError as why:
            raise SystemExit(f"Could not open {path}: {why.strerror}")

    try:
        vk = pproof.VerificationKey.from_json(slurp(args.verification_key_path))
    except (ValueError, KeyError, TypeError) as why:
        raise SystemExit(f"Could not deserialize verification key: {why}")
    try:
        proof = pproof.Proof.from_json(slurp(args.proof_path))
    except (ValueError, KeyError, TypeError) as why:
        raise SystemExit(f"Could not deserialize proof: {why}")
    print("Performing verification...")
    try:
        ok = backend.B200.verify(vk, proof)
    except ValueError as why:                         # malformed key / proof (the reference panics with the same reasons)
        raise SystemExit(f"Could not verify: {why}")
    print("PASSED" if ok else "FAILED")
    return 0


if __name__ == "__main__":
    sys.exit(main())
