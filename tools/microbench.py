#!/usr/bin/env python3
"""BASELINE.json config 5: MSM + NTT micro-benchmark sweep (BN254), GPU vs the ark-equivalent CPU port.

MSM points: the h_query of a GPU-made proving key (distinct non-infinity points tau^k*Z/delta*G), scalars
uniform below 2^252.  Times are device times from the library's CUDA-event stage timers (transfers excluded).
Test/bench infrastructure: the CPU column uses oracle/libzkoracle.so."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zokrates_b200._lib import Library, Context
from zokrates_b200 import synthetic

sizes = [int(a) for a in sys.argv[1:]] or [18, 20, 22]
cpu_max = int(os.environ.get("ZKB_CPU_MAX", "20"))
ctx = Context(0, 0, Library())
peak_mm = ctx.peak_probe(1, 4000)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
rs = np.random.RandomState(0x5EED0005 & 0x7FFFFFFF)
out = {"modmul_peak_per_s": peak_mm, "hbm_gbs_peak": peaks["hbm_gbs"], "rows": []}
oc = None
try:
    import __graft_entry__ as g
    from tests.oracle_c import OracleC
    oc = OracleC(g.build_oracle())
    out["cpu_threads"] = oc.threads()
except Exception as e:  # noqa
    print("no CPU oracle:", e)

lg_req = max(sizes)
lg_max = min(lg_req, int(os.environ.get("ZKB_POINTS_MAX_LOG", "24")))   # larger sizes tile this point set (timing only; noted per row)
r1cs, z = synthetic.make_layered(ctx, "bn128", (1 << lg_max) - 2)
h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
pk = ctx.setup(h, [11, 22, 33, 44, 5555, 3, 7])
ctx.r1cs_free(h)
m = r1cs.num_variables
off = 64 + 3 * 128 + 8 + 2 * 64 + 2 * 64 + (8 + m * 64) * 2 + 8 + m * 128 + 8   # start of h_query
hq = np.frombuffer(pk, dtype=np.uint8)[off:off + ((1 << lg_max) - 1) * 64]
del pk
for lg in sizes:
    n = (1 << lg) - 1 if lg >= lg_max else (1 << lg)
    reps = -(-n // ((1 << lg_max) - 1))
    pts = (np.tile(hq, reps)[:n * 64] if reps > 1 else hq[:n * 64]).tobytes()
    sc = rs.randint(0, 1 << 62, size=(n, 4)).astype(np.uint64); sc[:, 3] &= np.uint64((1 << 60) - 1)
    best = None
    for _ in range(3):
        res = ctx.msm(1, pts, sc)
        t = ctx.timings()
        if best is None or t["msm_exec"] < best["msm_exec"]:
            best = t
    row = {"log_n": lg, "n": n, "points_tiled": reps > 1, "msm_g1_ms": best["msm_plan"] + best["msm_exec"], "msm_plan_ms": best["msm_plan"],
           "msm_accum1_ms": best["accum1"], "msm_fq_mul_per_s": n * 16 * 10 / ((best["msm_plan"] + best["msm_exec"]) * 1e-3)}
    row["msm_frac_of_modmul_peak"] = row["msm_fq_mul_per_s"] / peak_mm
    x = rs.randint(0, 1 << 62, size=(1 << lg, 4)).astype(np.uint64); x[:, 3] &= np.uint64((1 << 60) - 1)
    tb = None
    for _ in range(3):
        y = ctx.ntt(x)
        t = ctx.timings()["ntt"]
        tb = t if tb is None else min(tb, t)
    row["ntt_ms"] = tb
    row["ntt_gbs"] = 64.0 * (1 << lg) / (tb * 1e-3) / 1e9
    row["ntt_frac_of_hbm"] = row["ntt_gbs"] / peaks["hbm_gbs"]
    row["ntt_fr_mul_per_s"] = (1 << lg) / 2 * lg / (tb * 1e-3)
    if oc and lg <= cpu_max:
        t0 = time.perf_counter(); ref = oc.msm(0, 1, pts, sc, 32); row["cpu_msm_g1_ms"] = (time.perf_counter() - t0) * 1e3
        assert ref == res, "MSM parity"
        t0 = time.perf_counter(); yr = oc.ntt(0, x); row["cpu_ntt_ms"] = (time.perf_counter() - t0) * 1e3
        assert np.array_equal(yr, y), "NTT parity"
    print(json.dumps(row), flush=True)
    out["rows"].append(row)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)
