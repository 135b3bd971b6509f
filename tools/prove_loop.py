#!/usr/bin/env python3
"""Profiling target: setup + N proofs of a synthetic 2^k circuit (use under ncu; numbers printed here are not bench values)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zokrates_b200._lib import Library, Context
from zokrates_b200 import synthetic
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dist = sys.argv[3] if len(sys.argv) > 3 else "uniform"
ctx = Context(0, 0, Library())
r1cs, z = synthetic.make_layered(ctx, "bn128", (1 << lg) - 2, distribution=dist)
h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
pk = ctx.setup(h, [11, 22, 33, 44, 5555, 3, 7])
pkh = ctx.pk_load(pk)
for i in range(reps):
    t = time.time(); ctx.prove(pkh, h, z, 5 + i, 7); print("prove", round((time.time() - t) * 1e3, 2), "ms", ctx.timings())
print("launches", ctx.launch_count())
