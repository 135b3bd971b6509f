"""Hostile bytes at the file-shaped entry points: a mutated program file, witness file or proving key must end in a status code
(ZkbError) or in a normal result, never in a crash — the reference panics on malformed input (`unwrap()`), the C ABI may not
abort the host process (include/zkb.h: "nothing throws or aborts across this boundary").  Deterministic mutations (seeded)."""
import random

import pytest

from tests.test_prog_native import solver_program
from zokrates_b200 import ir, zir
from zokrates_b200._lib import Context, ZkbError


def _mutations(data: bytes, rnd: random.Random, count: int):
    n = len(data)
    for _ in range(count):
        b = bytearray(data)
        kind = rnd.randrange(6)
        if kind == 0:                                   # flip a few bytes anywhere
            for _ in range(rnd.randrange(1, 4)):
                b[rnd.randrange(n)] ^= 1 << rnd.randrange(8)
        elif kind == 1:                                 # truncate
            b = b[:rnd.randrange(n)]
        elif kind == 2:                                 # corrupt the header / section table
            b[rnd.randrange(min(n, 100))] = rnd.randrange(256)
        elif kind == 3:                                 # overwrite a run with 0xff (CBOR breaks / huge lengths)
            i = rnd.randrange(n)
            b[i:i + rnd.randrange(1, 9)] = b"\xff" * rnd.randrange(1, 9)
        elif kind == 4:                                 # a length prefix that claims far more than the file holds
            i = rnd.randrange(n - 9)
            b[i] = 0x9b if rnd.random() < 0.5 else 0x5b  # array / byte string with a 64-bit length
            b[i + 1:i + 9] = (1 << 62).to_bytes(8, "big")
        else:                                           # deep nesting
            i = rnd.randrange(n)
            b[i:i] = b"\x81" * rnd.randrange(100, 400)
        yield bytes(b)


def test_program_file_mutations_emu(emu_lib):
    ctx = Context(0, 0, emu_lib)
    prog = solver_program("bn128", 8)
    data = zir.write_prog(prog)
    rnd = random.Random(2024)
    ok = bad = 0
    for blob in _mutations(data, rnd, 400):
        try:
            h = ctx.prog_load(blob)
        except ZkbError:
            bad += 1
            continue
        ok += 1
        try:
            info = ctx.prog_info(h)
            if info["schedulable"] and not info["unsupported_directives"] and info["arguments"] == 3:
                try:
                    ctx.prog_compute_witness(h, [201, 77, 5])
                except ZkbError:
                    pass
        finally:
            ctx.prog_free(h)
    assert bad > 100 and ok + bad == 400


def test_witness_file_mutations_emu(emu_lib):
    ctx = Context(0, 0, emu_lib)
    prog = solver_program("bn128", 8)
    h = ctx.prog_load(zir.write_prog(prog))
    wit = ir.Interpreter().execute(prog, [201, 77, 5]).write()
    rnd = random.Random(7)
    for blob in _mutations(wit, rnd, 200):
        try:
            ctx.prog_set_witness(h, blob)
        except ZkbError:
            pass
    ctx.prog_set_witness(h, wit)
    ctx.prog_free(h)


def test_proving_key_mutations_emu(emu_lib):
    from oracle import ark, gm17
    from oracle.ff import BN254
    from tests.util import rand_prog_pair
    from oracle import ir as oir
    c = BN254
    oprog, pprog, inputs = rand_prog_pair(c, 5, 1, 1, seed=3, curve_name="bn128")
    r1cs, _ = ark.synthesize(oprog, oir.execute(c, oprog, inputs))
    pk16 = ark.pk_serialize(c, ark.setup(c, r1cs, ark.Trapdoor(3, 5, 7, 11, 13)))
    pk17 = gm17.pk_serialize(c, gm17.setup(c, r1cs, gm17.Gm17Trapdoor(3, 5, 7, 11)))
    ctx = Context(0, 0, emu_lib)
    rnd = random.Random(11)
    for data, load, free in ((pk16, lambda b: ctx.pk_load(b), ctx.pk_free), (pk17, ctx.gm17_pk_load, ctx.gm17_pk_free)):
        for blob in _mutations(data, rnd, 150):
            try:
                h = load(blob)
            except ZkbError:
                continue
            free(h)
