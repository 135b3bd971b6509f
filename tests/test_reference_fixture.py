"""Proof VALUES against real zokrates_ark.

The reference's tests hold no golden proof (they assert `verify()` only, zokrates_ark/src/groth16.rs:113-160) and the
reference cannot be built in the authoring container, so the bit-exactness of whole proofs against arkworks rests on one
external step: `tools/make_reference_fixture.sh`, run where cargo exists, writes tests/golden/ref_<name>_<curve>/{out,
witness, proving.key.xz, verification.key, proof.json, meta.json}.  When those directories exist these tests hold
  * the CPU oracle  (python reader + RNG, C prover)                    [CPU tier]
  * the CUDA path   (file-level tool -> host mirror -> C ABI -> GPU)   [-m gpu]
to the reference's proof.json byte for byte.  Without them every test here x-fails with "parity unpinned" — visibly, not
silently."""
import glob
import json
import lzma
import os

import pytest

from oracle import ark
from oracle.ff import BLS12_381, BN254

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_*", "")))
CURVES = {"bn128": (0, BN254), "bls12_381": (1, BLS12_381)}
UNPINNED = ("parity unpinned: no real-zokrates_ark fixture under tests/golden/ref_*/ — run tools/make_reference_fixture.sh "
            "on a machine with cargo and commit its output")


def _load(d):
    meta = json.load(open(os.path.join(d, "meta.json")))
    pk = lzma.open(os.path.join(d, "proving.key.xz")).read()
    rd = lambda n: open(os.path.join(d, n), "rb").read()
    return meta, rd("out"), rd("witness"), pk, open(os.path.join(d, "proof.json")).read()


@pytest.mark.parametrize("d", FIXTURES or [None], ids=lambda d: os.path.basename(os.path.dirname(d)) if d else "absent")
def test_cpu_oracle_reproduces_the_reference_proof(d, oracle_c):
    if d is None:
        pytest.xfail(UNPINNED)
    from zokrates_b200 import ir, r1cs as pr1cs, zir
    meta, out, wit, pk, proof_json = _load(d)
    cid, c = CURVES[meta["curve"]]
    prog = zir.read_prog(out)
    witness = ir.Witness.read(wit, prog.curve)
    r1 = pr1cs.synthesize(prog)
    z = r1.assignment(witness)
    rng = ark.rng_from_entropy(meta["proof_entropy"])          # oracle-side Blake2b + ChaCha12 + ark Fr::rand
    r, s = ark.fr_rand(c, rng), ark.fr_rand(c, rng)
    raw, _ = oracle_c.prove(cid, pk, r1, z, r, s, c.fq_bytes)
    n = c.fq_bytes
    f = [int.from_bytes(raw[i * n:(i + 1) * n], "little") for i in range(8)]
    proof = ((f[0], f[1]), ((f[2], f[3]), (f[4], f[5])), (f[6], f[7]))
    inputs = [int(v) for v in prog.public_inputs_values(witness)]
    assert ark.tagged_proof_json(c, proof, inputs) == proof_json


@pytest.mark.gpu
@pytest.mark.parametrize("d", FIXTURES or [None], ids=lambda d: os.path.basename(os.path.dirname(d)) if d else "absent")
def test_gpu_path_reproduces_the_reference_proof(d, tmp_path, gpu_lib):
    if d is None:
        pytest.xfail(UNPINNED)
    from tools import zkb_generate_proof as tool
    meta, out, wit, pk, proof_json = _load(d)
    for name, data in (("out", out), ("witness", wit), ("proving.key", pk)):
        (tmp_path / name).write_bytes(data)
    rc = tool.main(["-i", str(tmp_path / "out"), "-w", str(tmp_path / "witness"), "-p", str(tmp_path / "proving.key"),
                    "-j", str(tmp_path / "proof.json"), "-e", meta["proof_entropy"]])
    assert rc == 0
    assert (tmp_path / "proof.json").read_text() == proof_json
