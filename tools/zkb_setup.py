#!/usr/bin/env python3
"""zkb-setup: `zokrates setup` for the b200 backend (/root/reference/zokrates_cli/src/ops/setup.rs:21-98,183-231, G16 only):
reads the compiled program, runs the circuit-specific Groth16 setup on the GPU (`zkb_groth16_setup`: QAP evaluation at tau and
fixed-base multiplications), writes `proving.key` in arkworks' uncompressed layout and `verification.key` as the reference's
TaggedVerificationKey JSON.  The trapdoor is drawn from the ChaCha RNG seeded as `get_rng_from_entropy` does (`-e`), or from
OS entropy.

    python tools/zkb_setup.py -i out -p proving.key -v verification.key [-e entropy]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="zkb-setup", description="Performs a trusted setup for a given constraint system")
    ap.add_argument("-i", "--input", default="out", metavar="FILE", help="Path of the binary")
    ap.add_argument("-p", "--proving-key-path", default="proving.key", metavar="FILE", help="Path of the generated proving key file")
    ap.add_argument("-v", "--verification-key-path", default="verification.key", metavar="FILE", help="Path of the generated verification key file")
    ap.add_argument("-b", "--backend", default="b200", choices=["b200"], help="Backend to use")
    ap.add_argument("-s", "--proving-scheme", default="g16", choices=["g16"], help="Proving scheme to use in the setup")
    ap.add_argument("-e", "--entropy", default=None, help="User provided randomness")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    from zokrates_b200 import backend, rng, zir
    try:
        with open(args.input, "rb") as f:
            prog = zir.read_prog(f.read())
    except OSError as why:
        raise SystemExit(f"Could not open {args.input}: {why.strerror}")
    except zir.ZirFormatError as why:
        raise SystemExit(str(why))
    print("Performing setup...")
    r = rng.get_rng_from_entropy(args.entropy) if args.entropy is not None else rng.StdRng.from_entropy()
    keypair = backend.B200.setup(prog, r, device=args.device)
    try:
        with open(args.verification_key_path, "w") as f:
            f.write(keypair.vk.to_tagged_json())
    except OSError as why:
        raise SystemExit(f"Could not create {args.verification_key_path}: {why.strerror}")
    print(f"Verification key written to '{args.verification_key_path}'")
    try:
        with open(args.proving_key_path, "wb") as f:
            f.write(keypair.pk)
    except OSError as why:
        raise SystemExit(f"Could not create {args.proving_key_path}: {why.strerror}")
    print(f"Proving key written to '{args.proving_key_path}'")
    print("Setup completed")
    return 0


if __name__ == "__main__":
    sys.exit(main())
