#!/usr/bin/env python3
"""First-contact GPU check (run under gpurun): parity of the product library against the python oracle on
small inputs, then timings of the building blocks.  Test infrastructure (imports oracle/)."""
import json, os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zokrates_b200._lib import Library, Context, fr_array, fr_from_array
from oracle.ff import BN254, BLS12_381, g1_group, g2_group
from oracle import ark
from oracle.ir import Prog as OProg, Constraint as OC, var_new, var_public, execute

lib = Library(os.environ.get("ZKB200_LIB"))
random.seed(11)
out = {}


def csr(rows, r):
    rowptr = [0]; col = []; val = []
    for row in rows:
        for cidx, k in row:
            col.append(cidx); val.append(k % r)
        rowptr.append(len(col))
    return np.array(rowptr, dtype=np.uint64), np.array(col, dtype=np.uint32), fr_array(val)


def rand_prog(c, ncons, npub, npriv):
    args = [(var_new(i), i >= npub) for i in range(npub + npriv)]
    nxt = npub + npriv
    stmts = []
    avail = [v for v, _ in args]
    for j in range(ncons):
        def lc():
            return [(random.choice(avail + [0]), random.choice([1, 1, 2, c.r - 1, random.randrange(c.r)]))
                    for _ in range(random.choice([1, 1, 2, 3]))]
        o = var_new(nxt); nxt += 1
        if j == ncons - 1:
            o = var_public(0)
        stmts.append(OC(lc(), lc(), [(o, 1)]))
        avail.append(o)
    return OProg(args, 1, stmts)


def check_curve(cid, c, quick=False):
    ctx = Context(cid, 0, lib)
    a = [random.randrange(c.r) for _ in range(1000)]; b = [random.randrange(c.r) for _ in range(1000)]
    assert fr_from_array(ctx.field_op(0, 0, fr_array(a), fr_array(b))) == [x * y % c.r for x, y in zip(a, b)]
    nq = c.fq_bytes // 8
    a = [0, 1, c.p - 1] + [random.randrange(c.p) for _ in range(500)]; b = [c.p - 1, 0, c.p - 1] + [random.randrange(c.p) for _ in range(500)]
    for op, f in ((0, lambda x, y: x * y % c.p), (1, lambda x, y: (x + y) % c.p), (2, lambda x, y: (x - y) % c.p)):
        assert fr_from_array(ctx.field_op(1, op, fr_array(a, nq), fr_array(b, nq))) == [f(x, y) for x, y in zip(a, b)], op
    assert fr_from_array(ctx.field_op(1, 3, fr_array(a[2:50], nq), None)) == [pow(x, -1, c.p) for x in a[2:50]]
    print(c.name, "field ok", flush=True)
    for logn in (0, 1, 3, 4, 6, 9, 10):
        n = 1 << logn
        x = [random.randrange(c.r) for _ in range(n)]
        d = ark.Domain(c, n)
        assert fr_from_array(ctx.ntt(fr_array(x))) == d.fft(x), logn
        assert fr_from_array(ctx.ntt(fr_array(x), inverse=True, coset=True)) == d.coset_ifft(x), logn
    print(c.name, "ntt ok", flush=True)
    G1, G2 = g1_group(c), g2_group(c)
    for n in (0, 1, 5, 200):
        pts = [G1.mul(c.g1, random.randrange(1, c.r)) for _ in range(n)]
        sc = [random.choice([0, 1, 2, c.r - 1, random.randrange(c.r), random.randrange(1 << 20)]) for _ in range(n)]
        if n > 3:
            pts[2] = None; pts[3] = pts[1]; sc[3] = sc[1]
        got = ctx.msm(1, b"".join(ark.ser_g1(c, p) for p in pts), fr_array(sc))
        assert got == ark.ser_g1(c, G1.msm_naive(pts, sc)), ("g1", n)
    for n in (0, 3, 40):
        pts = [G2.mul(c.g2, random.randrange(1, c.r)) for _ in range(n)]
        sc = [random.choice([0, 1, random.randrange(c.r)]) for _ in range(n)]
        assert ctx.msm(2, b"".join(ark.ser_g2(c, p) for p in pts), fr_array(sc)) == ark.ser_g2(c, G2.msm_naive(pts, sc)), ("g2", n)
    print(c.name, "msm small ok", flush=True)
    # few distinct points repeated: exercises big buckets, same-point doubling, level >= 2 reduction
    base = [G1.mul(c.g1, random.randrange(1, c.r)) for _ in range(16)]
    for n, dist in ((1 << 12, "uniform"), (1 << 14, "bits"), (1 << 16, "uniform")):
        if quick and n > 1 << 14:
            continue
        sc = [random.randrange(c.r) if dist == "uniform" else random.choice([0, 1, 1, 1, random.randrange(256)]) for _ in range(n)]
        ptsb = b"".join(ark.ser_g1(c, base[i % 16]) for i in range(n))
        t = time.time()
        got = ctx.msm(1, ptsb, fr_array(sc))
        dt = time.time() - t
        sums = [sum(sc[j::16]) % c.r for j in range(16)]
        assert got == ark.ser_g1(c, G1.msm_naive(base, sums)), ("g1 big", n, dist)
        print(c.name, "msm", n, dist, "ok", round(dt, 3), ctx.timings(), flush=True)
    for (ncons, npub, npriv) in ((1, 1, 1), (13, 0, 3), (100, 1, 2)):
        prog = rand_prog(c, ncons, npub, npriv)
        w = execute(c, prog, [random.randrange(c.r) for _ in range(npub + npriv)])
        r1cs, z = ark.synthesize(prog, w)
        h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness,
                          [csr(r1cs.a, c.r), csr(r1cs.b, c.r), csr(r1cs.c, c.r)])
        n = ark.Domain(c, r1cs.num_constraints + r1cs.num_instance).n
        assert fr_from_array(ctx.witness_map(h, fr_array(z), n)) == ark.witness_map(c, r1cs, z), "witness_map"
        td = ark.Trapdoor(*[random.randrange(1, c.r) for _ in range(7)])
        pkb = ctx.setup(h, [td.alpha, td.beta, td.gamma, td.delta, td.tau, td.g1_k, td.g2_k])
        if ncons <= 13:
            assert pkb == ark.pk_serialize(c, ark.setup(c, r1cs, td)), "setup"
        pkh = ctx.pk_load(pkb)
        rr, ss = random.randrange(c.r), random.randrange(c.r)
        proof = ctx.prove(pkh, h, fr_array(z), rr, ss)
        exp = ark.trapdoor_expected_proof(c, r1cs, td, z, rr, ss)
        assert proof == ark.ser_g1(c, exp[0]) + ark.ser_g2(c, exp[1]) + ark.ser_g1(c, exp[2]), "proof"
        parts = [ctx.prove_partial(ctx.pk_load(pkb, rank, 3), h, fr_array(z)) for rank in range(3)]
        assert ctx.finalize(pkh, np.concatenate(parts), 3, rr, ss) == proof, "sharded"
        print(c.name, ncons, "prove ok", ctx.timings(), flush=True)
    return ctx


t0 = time.time()
ctx = check_curve(0, BN254)
out["imad_per_s"] = ctx.peak_probe(0, 40000)
out["modmul_per_s"] = ctx.peak_probe(1, 4000)
print("peaks", out, flush=True)
check_curve(1, BLS12_381, quick=True)
print("ALL OK", round(time.time() - t0, 1), "s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/selftest.json", "w"))
