"""sha256packed as a PROGRAM with solver directives (`Bits`, `Xor`, `ShaCh`, `ShaAndXorAndXorAnd`) through the native front door:
witness generation must run ~28 k solver directives between ~55 k constraints, level by level on the device.  Pinned to the
reference's KAT (zokrates_stdlib/tests/tests/hashes/sha256/512bitPacked.json:5-16: inputs 0, 0, 0, 5) and to the host mirror of
the interpreter, byte for byte.  CPU tier: host-emulation build; -m gpu: libzkb200.so."""
import hashlib

import pytest

from zokrates_b200 import ir, sha256_circuit, zir
from zokrates_b200._lib import Context

KAT = (263561599766550617289250058199814760685, 65303172752238645975888084098459749904)


@pytest.fixture(scope="module")
def program():
    prog = sha256_circuit.make_prog("bn128")
    return prog, zir.write_prog(prog)


def _check(lib, program, inputs):
    prog, data = program
    digest = hashlib.sha256(b"".join(int(v).to_bytes(16, "big") for v in inputs)).digest()
    want = [int.from_bytes(digest[:16], "big"), int.from_bytes(digest[16:], "big")]
    ref = ir.Interpreter().execute(prog, inputs)
    assert ref.return_values() == want
    ctx = Context(0, 0, lib)
    h = ctx.prog_load(data)
    try:
        info = ctx.prog_info(h)
        assert info["directives"] > 20000 and info["unsupported_directives"] == 0 and info["schedulable"] == 1
        wit = ctx.prog_compute_witness(h, inputs)
        assert wit == ref.write()
        assert ctx.prog_public_inputs(h) == want
    finally:
        ctx.prog_free(h)
    return want


def test_sha256_program_kat_emu(program, emu_lib):
    assert tuple(_check(emu_lib, program, [0, 0, 0, 5])) == KAT


@pytest.mark.gpu
def test_sha256_program_gpu(program, gpu_lib):
    assert tuple(_check(gpu_lib, program, [0, 0, 0, 5])) == KAT
    _check(gpu_lib, program, [2 ** 128 - 1, 12345678901234567890, 0, 2 ** 127 + 99])
