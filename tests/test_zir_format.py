"""The compiled-program file (`out`): header layout and serde_cbor encoding restated from
/root/reference/zokrates_ast/src/ir/serialize.rs.  CPU only.  The reference holds no golden `out`
file; what it does pin — magic, version, the two curve ids, the empty-program round trip of `ser_deser_v2` (:400-426) — is
checked here, together with hand-assembled CBOR that follows serde_cbor's documented struct/enum encoding."""
import io
import struct

import pytest

from zokrates_b200 import ir, zir
from zokrates_b200.ir import Constraint, Directive, LinComb, Parameter, Prog, QuadComb, Variable


def test_empty_program_round_trip_both_curves():
    # ser_deser_v2: Prog::default() for Bn128 and Bls12_381 survives serialize -> deserialize
    for curve, cid in (("bn128", "b4f7b5bd"), ("bls12_381", "40d8c1f9")):
        data = zir.write_prog(Prog(curve=curve))
        assert data[:4] == b"ZOK\x00" and data[4:8] == bytes([3, 0, 0, 0]) and data[8:12].hex() == cid
        p = zir.read_prog(data)
        assert p.curve == curve and p.arguments == [] and p.statements == [] and p.return_count == 0


def test_header_layout_and_quirks():
    x, y = Variable.new(0), Variable.new(1)
    prog = Prog([Parameter.private_(x), Parameter.public(y)], 1, [ir.constraint(x, x, y), ir.definition(Variable.public(0), y)])
    data = zir.write_prog(prog)
    name, n_cons, n_ret, sec = zir.read_header(data)
    assert (name, n_cons, n_ret) == ("bn128", 2, 1)
    assert [s[0] for s in sec] == [1, 2, 3, 3]               # module map written with the Solvers id (serialize.rs:252)
    assert sec[0][1] == 120                                   # size_of::<ProgHeader>() reserved, 100 bytes used
    assert data[100:120] == b"\x00" * 20
    for (_, off, ln), (_, off2, _) in zip(sec, sec[1:]):
        assert off + ln == off2                               # sections are contiguous
    assert sec[3][1] + sec[3][2] == len(data)


def test_cbor_encoding_of_a_constraint_is_serde_cbor_shaped():
    x = Variable.new(0)
    data = zir.write_prog(Prog([Parameter.private_(x)], 0, [ir.constraint(x, x, Variable.one())]))
    _, _, _, sec = zir.read_header(data)
    params = data[sec[0][1]:sec[0][1] + sec[0][2]]
    # Vec<Parameter{span: None, id: Variable{id: 1}, private: true}>
    assert params == (b"\x81\xa3" + b"\x64span\xf6" + b"\x62id\xa1\x62id\x01" + b"\x67private\xf5")
    st = data[sec[1][1]:sec[1][1] + sec[1][2]]
    one = (1).to_bytes(32, "little")
    lc = lambda vid: b"\xa2\x64span\xf6\x65value\x81\x82\xa1\x62id" + bytes([vid]) + b"\x58\x20" + one
    expect = (b"\xa1\x6aConstraint\xa4" + b"\x64span\xf6" + b"\x64quad\xa3\x64span\xf6\x64left" + lc(1) + b"\x65right" + lc(1)
              + b"\x63lin" + lc(0) + b"\x65error\xf6")
    assert st == expect


def test_program_round_trip_with_directives_and_negative_ids():
    x, b0, b1, out = Variable.new(0), Variable.new(1), Variable.new(2), Variable.public(0)
    r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    stmts = [
        Directive([QuadComb(LinComb.one(), LinComb.from_var(x))], [b0, b1], "Bits", 2),
        Constraint(QuadComb(LinComb.from_var(b0), LinComb.from_var(b0)), LinComb.from_var(b0), "Bitness"),
        Constraint(QuadComb(LinComb.from_var(b1), LinComb.from_var(b1)), LinComb.from_var(b1)),
        Constraint(QuadComb(LinComb([(b0, 2), (b1, 1)]), LinComb.one()), LinComb.from_var(x)),
        Directive([QuadComb(LinComb.one(), LinComb.from_var(b0)), QuadComb(LinComb.one(), LinComb.from_var(b1))], [Variable.new(3)], "Xor"),
        Constraint(QuadComb(LinComb([(x, r - 1)]), LinComb.one()), LinComb([(x, r - 1)])),
        Constraint(QuadComb(LinComb.from_var(x), LinComb.one()), LinComb.from_var(out)),
    ]
    prog = Prog([Parameter.public(x)], 1, stmts)
    back = zir.read_prog(zir.write_prog(prog))
    assert back.arguments == prog.arguments and back.return_count == 1
    assert [type(s) for s in back.statements] == [type(s) for s in stmts]
    for a, b in zip(back.statements, stmts):
        if isinstance(a, Constraint):
            assert (a.quad, a.lin, a.error) == (b.quad, b.lin, b.error)
        else:
            assert (a.solver, a.arg, a.outputs, a.inputs) == (b.solver, b.arg, b.outputs, b.inputs)
    # the interpreter gives the same witness on the decoded program
    w0 = ir.Interpreter().execute(prog, [2]).values
    w1 = ir.Interpreter().execute(back, [2]).values
    assert w0 == w1 and w0[out] == 2


def test_reader_accepts_what_serde_cbor_accepts():
    # indefinite-length containers, a tag in front of a value, a field element given as an array of integers
    body = io.BytesIO()
    body.write(b"\x00" * 120)
    params_off = body.tell(); body.write(b"\x9f\xff")                     # [_ ] empty indefinite array
    cons_off = body.tell()
    one_arr = b"\x98\x20" + b"\x01" + b"\x00" * 31                         # array(32) of small ints
    lc = b"\xbf\x64span\xf6\x65value\x9f\x82\xa1\x62id\x20" + one_arr + b"\xff\xff"   # id -1 (~out_0), indefinite map/array
    body.write(b"\xd9\xd9\xf7" + b"\xa1\x6aConstraint\xa4\x64span\xf6\x64quad\xa3\x64span\xf6\x64left" + lc + b"\x65right" + lc
               + b"\x63lin" + lc + b"\x65error\xa1\x6fSourceAssertion\xa0")
    sol_off = body.tell(); body.write(b"\x80")
    mod_off = body.tell(); body.write(b"\xa1\x67modules\xa0")
    end = body.tell()
    head = b"ZOK\x00" + bytes([3, 0, 0, 0]) + bytes.fromhex("b4f7b5bd") + struct.pack("<II", 1, 1)
    for ty, a, b in ((1, params_off, cons_off), (2, cons_off, sol_off), (3, sol_off, mod_off), (3, mod_off, end)):
        head += struct.pack("<IQQ", ty, a, b - a)
    data = bytearray(body.getvalue()); data[:100] = head
    p = zir.read_prog(bytes(data))
    (s,) = p.statements
    assert s.lin.value == [(Variable.public(0), 1)] and s.error == "SourceAssertion"


def test_rejections():
    good = zir.write_prog(Prog([Parameter.public(Variable.new(0))], 0, [ir.constraint(Variable.new(0), Variable.new(0), Variable.new(0))]))
    for mutate, msg in ((lambda d: d.__setitem__(0, 0x58), "Invalid magic number"), (lambda d: d.__setitem__(4, 2), "Invalid file version"),
                        (lambda d: d.__setitem__(8, 0), "Unknown curve identifier"), (lambda d: d.__setitem__(20, 9), "invalid section type"),
                        (lambda d: d.__setitem__(12, 7), "constraint count")):
        d = bytearray(good); mutate(d)
        with pytest.raises(zir.ZirFormatError, match=msg):
            zir.read_prog(bytes(d))
    with pytest.raises(zir.ZirFormatError, match="Invalid header"):
        zir.read_prog(good[:50])
    d = bytearray(good)
    _, _, _, sec = zir.read_header(good)
    k = good.index((1).to_bytes(32, "little"), sec[1][1])
    d[k:k + 32] = b"\xff" * 32                                # coefficient >= modulus
    with pytest.raises(zir.ZirFormatError, match="non-canonical"):
        zir.read_prog(bytes(d))


def test_file_level_pipeline_on_the_emulated_engine(tmp_path, emu_lib, monkeypatch):
    """`out` + arguments -> zkb-compute-witness -> `witness` (-> .json, .wtns) -> zkb-generate-proof -> `proof.json`, the two
    file-level tools chained like `zokrates compute-witness` / `generate-proof`; kernels run in the host-emulation library
    (ZKB200_LIB), results compared with the host interpreter and the python oracle's proof."""
    import io, json
    from oracle import ark, ir as oir
    from oracle.ff import BN254
    from tools import zkb_compute_witness, zkb_generate_proof
    from zokrates_b200 import _lib, backend, circom
    monkeypatch.setenv("ZKB200_LIB", emu_lib.path)
    monkeypatch.setattr(_lib, "_default", None)
    monkeypatch.setattr(backend, "_contexts", {})
    a_, b_ = Variable.new(0), Variable.new(1)
    prog = Prog([Parameter.private_(a_), Parameter.public(b_)], 0, [ir.constraint(a_, a_, b_)], "bn128")
    (tmp_path / "out").write_bytes(zir.write_prog(prog))
    rc = zkb_compute_witness.main(["-i", str(tmp_path / "out"), "-o", str(tmp_path / "witness"), "-a", "337", "113569", "--json",
                                   "--circom-witness", str(tmp_path / "out.wtns")])
    assert rc == 0
    ref = ir.Interpreter().execute(prog, [337, 113569])
    assert (tmp_path / "witness").read_bytes() == ref.write()
    assert (tmp_path / "witness.json").read_text() == ref.write_json()
    assert (tmp_path / "out.wtns").read_bytes() == circom.write_witness(ref, [b_])
    with pytest.raises(SystemExit, match="Execution failed"):
        zkb_compute_witness.main(["-i", str(tmp_path / "out"), "-o", str(tmp_path / "w2"), "-a", "3", "10"])
    with pytest.raises(SystemExit, match="Could not parse argument"):
        zkb_compute_witness.main(["-i", str(tmp_path / "out"), "-o", str(tmp_path / "w2"), "-a", "x"])
    from tools import zkb_setup, zkb_verify
    rc = zkb_setup.main(["-i", str(tmp_path / "out"), "-p", str(tmp_path / "proving.key"), "-v", str(tmp_path / "verification.key"),
                         "-e", "ceremony"])
    assert rc == 0
    from zokrates_b200 import rng as prng2, proof as pproof2
    c_prod = backend._curve("bn128")
    r_setup = prng2.get_rng_from_entropy("ceremony")
    td = [prng2.fr_rand(c_prod, r_setup) for _ in range(7)]       # alpha, beta, gamma, delta, tau and the two generator scalars
    kp = backend.B200.setup(prog, td)
    assert (tmp_path / "proving.key").read_bytes() == kp.pk
    assert (tmp_path / "verification.key").read_text() == kp.vk.to_tagged_json()
    assert pproof2.VerificationKey.from_json(kp.vk.to_tagged_json()) == kp.vk
    rc = zkb_generate_proof.main(["-i", str(tmp_path / "out"), "-w", str(tmp_path / "witness"), "-p", str(tmp_path / "proving.key"),
                                  "-j", str(tmp_path / "proof.json"), "-e", "pipeline"])
    assert rc == 0
    c = BN254
    oprog = oir.Prog([(a_.id, True), (b_.id, False)], 0, [oir.Constraint([(a_.id, 1)], [(a_.id, 1)], [(b_.id, 1)])])
    ow = oir.execute(c, oprog, [337, 113569])
    r1cs_o, z = ark.synthesize(oprog, ow)
    orng = ark.rng_from_entropy("pipeline")
    r, s = ark.fr_rand(c, orng), ark.fr_rand(c, orng)
    exp = ark.trapdoor_expected_proof(c, r1cs_o, ark.Trapdoor(*td), z, r, s)
    assert (tmp_path / "proof.json").read_text() == ark.tagged_proof_json(c, exp, [113569])
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        assert zkb_verify.main(["-j", str(tmp_path / "proof.json"), "-v", str(tmp_path / "verification.key")]) == 0
    assert buf.getvalue().splitlines() == ["Performing verification...", "PASSED"]
    tampered = json.loads((tmp_path / "proof.json").read_text())
    tampered["inputs"][0] = "0x" + (113570).to_bytes(32, "big").hex()
    (tmp_path / "bad.json").write_text(json.dumps(tampered))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        zkb_verify.main(["-j", str(tmp_path / "bad.json"), "-v", str(tmp_path / "verification.key")])
    assert buf.getvalue().splitlines()[-1] == "FAILED"
