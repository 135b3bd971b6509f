#!/usr/bin/env python3
"""Stage timings of setup / pk_load / prove on synthetic circuits (run under gpurun)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zokrates_b200._lib import Library, Context
from zokrates_b200 import synthetic

lib = Library(os.environ.get("ZKB200_LIB"))
curve = os.environ.get("ZKB_CURVE", "bn128")
ctx = Context(0 if curve == "bn128" else 1, 0, lib)
sizes = [int(a) for a in sys.argv[1:]] or [16, 18, 20]
res = {}
for lg in sizes:
    for dist in ("uniform", "bits"):
        t = time.time()
        r1cs, z = synthetic.make_layered(ctx, curve, (1 << lg) - 2, distribution=dist)
        tgen = time.time() - t
        h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
        t = time.time()
        pk = ctx.setup(h, [11, 22, 33, 44, 5555, 3, 7])
        tsetup = time.time() - t
        st = ctx.timings()
        t = time.time()
        pkh = ctx.pk_load(pk)
        tload = time.time() - t
        best = None
        for it in range(3):
            t = time.time()
            proof = ctx.prove(pkh, h, z, 1234567 + it, 7654321)
            dt = time.time() - t
            tm = ctx.timings()
            if best is None or dt < best[0]:
                best = (dt, tm)
        key = f"{curve}/2^{lg}/{dist}"
        res[key] = {"gen_s": round(tgen, 2), "setup_s": round(tsetup, 3), "setup_stages_ms": st, "pk_bytes": len(pk),
                    "pk_load_s": round(tload, 3), "prove_wall_ms": round(best[0] * 1e3, 3), "stages_ms": best[1],
                    "constraints_per_s": round(r1cs.num_constraints / best[0])}
        print(key, json.dumps(res[key]), flush=True)
        ctx.pk_free(pkh); ctx.r1cs_free(h)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/profile_stages_{curve}.json", "w"), indent=1)
