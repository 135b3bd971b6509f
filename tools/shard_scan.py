#!/usr/bin/env python3
"""Per-rank cost of the index-sharded prover, measured on ONE GPU: load the proving key as rank 0 of W for W = 1, 2, 4, 8
and time prove_partial (the replicated witness map + this rank's MSM shards).  Tells what bounds strong scaling
before spending multi-GPU box time.  Not a bench value (no NCCL, one rank)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zokrates_b200._lib import Library, Context
from zokrates_b200 import synthetic

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
worlds = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
ctx = Context(0, 0, Library())
r1cs, z = synthetic.make_layered(ctx, "bn128", (1 << lg) - 2)
h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
pk = ctx.setup(h, [11, 22, 33, 44, 5555, 3, 7])
ctx.set_assignment(h, z)
rows = []
for w in worlds:
    pkh = ctx.pk_load(pk, w - 1, w)
    best, stages = None, None
    for i in range(6):
        t = time.perf_counter()
        ctx.prove_partial(pkh, h, None)
        dt = (time.perf_counter() - t) * 1e3
        if i >= 2 and (best is None or dt < best):
            best, stages = dt, ctx.timings()
    ctx.pk_free(pkh)
    rows.append({"world": w, "ms": best, "stages": stages})
    print(json.dumps(rows[-1]))
