#!/usr/bin/env python3
"""bench.py — Groth16 constraints/sec on synthetic R1CS, BN254, N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo (libzkb200.so, sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores

Workload (BASELINE.json config 3): synthetic R1CS with 2^20 - 2 constraints, one public input, so the
evaluation domain is exactly 2^20; uniform 252-bit witness (MSM worst case).  A "step" is one proof:
witness_map (3 SpMV + 7 NTT) + 4 G1 MSM + 1 G2 MSM + final combination.
  value : whole-job constraints/s with z, the CSR matrices and the proving key resident in HBM
  e2e   : the same through the C-ABI call a `zokrates_b200` Rust shim makes (zkb_groth16_prove): z in pinned
          host memory, H2D of z and D2H of the window sums / proof inside the timed region
Multi-GPU: every MSM is sharded by index range over the ranks (no data-path collective); the 5 partial
sums per rank are all-gathered (NCCL) and rank 0 finishes the proof; the witness map is replicated for N <= 2 and
its three chains are computed once each and broadcast (NCCL over NVLink) for N >= 3.
The reference arm times oracle/libzkoracle.so — the C restatement of ark's prover (the reference is
Rust + un-vendored arkworks crates and cannot be built here) — on all host threads, on a
bounded sample of the same circuit family.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 20
TRAPDOOR = [0x1111, 0x2222, 0x3333, 0x4444, 0x123456789ABCDEF, 3, 7]
SAMPLE_LOG_N = 16          # bounded sample for the CPU arms


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md).  Started before the warm-up (nvidia-smi takes a few
    hundred ms to come up, longer than a short timed region); every sample carries the host time it was read at and
    only those inside the marked timed windows are reported (all load samples if a window caught none)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_ms=50):
        self.index, self.period_ms, self.proc, self.lines, self.windows = index, period_ms, None, [], []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line))

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc:
            time.sleep(2.5 * self.period_ms / 1e3)
            self.proc.terminate()
        rows = []
        for ts, line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                rows.append((ts, float(p[1]), float(p[2]), float(p[3]), p[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if any(a <= r[0] <= b + self.period_ms / 1e3 for a, b in self.windows)]
        scope = "timed region"
        if not inside:                       # fall back to the samples taken under load (warm-up + timed steps)
            inside = [r for r in rows if r[3] > 250.0] or rows
            scope = "warm-up + timed steps (no sample fell inside the timed windows)"
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope,
                "power_w_max": max((r[3] for r in inside), default=None)}


def load_oracle():
    """CPU checker / baseline (oracle/libzkoracle.so).  Only the cpu_baseline leg and --impl reference use it."""
    import __graft_entry__ as g
    from tests.oracle_c import OracleC
    return OracleC(g.build_oracle())


def cpu_prove_sample(steps, warmup):
    """The reference's CPU algorithm on a bounded sample: 2^16 - 2 constraints of the same circuit family."""
    from zokrates_b200 import synthetic
    oc = load_oracle()
    n_cons = (1 << SAMPLE_LOG_N) - 2
    r1cs, z = synthetic.make("bn128", n_cons, distribution="uniform")
    pk = oc.setup(0, r1cs, TRAPDOOR)
    times = []
    proof = None
    for i in range(warmup + steps):
        t = time.perf_counter()
        proof, stage = oc.prove(0, pk, r1cs, z, 1234567, 7654321, 32)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    return {"n_cons": n_cons, "times": times, "threads": oc.threads(), "r1cs": r1cs, "z": z, "pk": pk, "proof": proof,
            "stage_s": [float(x) for x in stage]}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    res = cpu_prove_sample(args.steps, min(args.warmup, 1))
    total = sum(res["times"])
    value = res["n_cons"] * len(res["times"]) / total
    line = {
        "impl": "reference", "metric": "groth16_constraints_per_sec", "value": value, "unit": "constraints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(res["times"]),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256-montgomery", "data": "synthetic",
        "config": {"workload": "synthetic-r1cs-2^20-bn128-groth16", "sample": f"2^{SAMPLE_LOG_N}-2 constraints of the same generator",
                   "curve": "bn128"},
        "cpu_baseline": {"value": value, "unit": "constraints/s", "cores": res["threads"], "kind": "port",
                         "sample": f"{len(res['times'])} proofs of a 2^{SAMPLE_LOG_N}-2 constraint synthetic circuit (ark-equivalent C port, "
                                   "MSM parallel over windows only as in ark 0.3.0)"},
        "e2e": {"value": value, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_s": res["stage_s"],
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=LOG_N)
    ap.add_argument("--witness", default="uniform", choices=["uniform", "bits"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line (the JSON): libraries that print on fd 1 (NCCL's version banner, whatever
    # NCCL_DEBUG / nccl.conf say) go to stderr for the duration of the run
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from zokrates_b200 import synthetic
    from zokrates_b200._lib import Context, Library

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libzkb200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and not os.environ.get("ZKB_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL prints its banner on stdout: keep stdout to the one JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    lib = Library()
    ctx = Context(0, local, lib)
    n_cons = (1 << args.log_n) - 2

    # -- untimed preparation: circuit, witness (batched field ops on the GPU), setup, resident key shard
    t_prep = time.perf_counter()
    r1cs, z = synthetic.make_layered(ctx, "bn128", n_cons, distribution=args.witness)
    r1cs_h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
    pk = ctx.setup(r1cs_h, TRAPDOOR)
    setup_ms = ctx.timings()
    pk_h = ctx.pk_load(pk, rank, world)
    pk_bytes = len(pk)
    del pk
    z_pinned = torch.from_numpy(z).pin_memory()
    z_host = z_pinned.numpy()
    ctx.set_assignment(r1cs_h, z_host)
    prep_s = time.perf_counter() - t_prep
    r_s = (1234567, 7654321)
    partial_bytes = ctx.partial_bytes

    def gather_and_finish(partial):
        """5 partial sums per rank -> all ranks (NCCL all_gather of a few hundred bytes) -> rank 0 finishes."""
        if world == 1:
            return ctx.finalize(pk_h, partial, 1, *r_s)
        from zokrates_b200.distributed import gather_partials
        allp = gather_partials(partial, device=torch.device("cuda", local))
        return ctx.finalize(pk_h, allp, world, *r_s) if rank == 0 else None

    last_stage = {}

    # three or more ranks: the witness map's three chains are computed once each and broadcast over NVLink
    # (zkb_groth16_prove_begin / _end, zokrates_b200/distributed.py); ZKB_WM_SHARE=0 keeps it replicated
    share_wm = world >= 3 and os.environ.get("ZKB_WM_SHARE", "1") != "0"

    def partial_step(z_arg):
        if rank == 0:
            ctx.finalize_prepare(pk_h, *r_s)           # r*delta1 ... s*delta2 on host threads underneath the kernels
        if share_wm:
            from zokrates_b200.distributed import prove_partial_shared_wm
            partial = prove_partial_shared_wm(ctx, pk_h, r1cs_h, z_arg, device=torch.device("cuda", local))
        else:
            partial = ctx.prove_partial(pk_h, r1cs_h, z_arg)
        last_stage.update(ctx.timings())
        return gather_and_finish(partial)

    def step_resident():
        if world == 1:
            out = ctx.prove_resident(pk_h, r1cs_h, *r_s)
            last_stage.update(ctx.timings())
            return out
        return partial_step(None)

    def step_e2e():
        if world == 1:
            out = ctx.prove(pk_h, r1cs_h, z_host, *r_s)
            last_stage.update(ctx.timings())
            return out
        return partial_step(z_host)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step_fn, steps):
        """K steps bracketed by barrier + synchronize; device time via CUDA events on the current stream (each
        step ends with a stream sync inside the library, so the event pair brackets all device work), max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ctx.launch_count()
        t0 = time.perf_counter()
        e0.record()
        stage = {}
        proof = None
        for _ in range(steps):
            proof = step_fn()
            for k, v in last_stage.items():
                stage[k] = stage.get(k, 0.0) + v
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) / 1e3
        elapsed = max(wall, dev_s)
        if world > 1:
            tt = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            elapsed = float(tt.item())
        return elapsed, {k: v / steps for k, v in stage.items()}, ctx.launch_count() - launches0, proof

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    tw0 = time.perf_counter()
    t_res, stage_res, launches, proof = timed(step_resident, args.steps)
    sampler.mark(tw0, time.perf_counter())
    for _ in range(2):
        step_e2e()
    tw0 = time.perf_counter()
    t_e2e, stage_e2e, _, proof2 = timed(step_e2e, args.steps)
    sampler.mark(tw0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0 and proof != proof2:
        raise SystemExit("resident and e2e proofs differ")

    value = n_cons * args.steps / t_res
    e2e_value = n_cons * args.steps / t_e2e

    # -- roofline of the dominant kernel (msm_accum1: bucket accumulation of the three full-size G1 MSMs)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # roofline kernel: the bucket accumulation of the h_query MSM (no infinity points, uniform scalars, so the
    # canonical n*W*10 count is not flattered by the infinity-skipping views used for a/b1/b2)
    n_pairs = (r1cs.domain_size - 1) // world            # (scalar, point) pairs one accum1 launch of msm_h processes
    acc_ms = stage_res.get("accum1_g1_h", 0.0)
    roofline = roofline_mm = None
    if rank == 0 and acc_ms > 0:
        modmul_peak = ctx.peak_probe(1, 4000)
        imad_peak = ctx.peak_probe(0, 40000)
        alg_bytes = n_pairs * 96.0                       # 32 B scalar + 64 B affine point per pair
        traffic = None                                   # DRAM bytes per launch from the committed ncu --set full capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_msm_accum1<Fq> (h_query MSM)"]
            traffic = tr["dram_bytes"] / world
        except (OSError, KeyError, ValueError):
            pass
        alg_muls = n_pairs * 16 * 10.0                   # canonical: W = 16 windows x 10 Fq-mul per mixed add
        roofline = {"kernel": "k_msm_accum1<Fq> (h_query MSM)", "bound": "hbm", "achieved": alg_bytes / (acc_ms * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (acc_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                    "traffic_note": "ncu dram__bytes_read+write of this kernel (profiles/r01_ncu_accum1_final.md): the gathers go to 5 GB of "
                                    "HBM-resident window tables 2^(cw)P, a deliberate bytes-for-multiplications trade (14 instead of 16 mixed additions per scalar, one bucket set)",
                    "peak_source": hbm_src, "avg_launch_ms": acc_ms,
                    "note": "the kernel is integer-multiply bound, not HBM bound (230 MAD/B): see roofline_modmul",
                    "binding_resource": "int32-mad", "frac_of_binding_resource": alg_muls / (acc_ms * 1e-3) / modmul_peak}
        roofline_mm = {"kernel": "k_msm_accum1<Fq> (h_query MSM)", "bound": "int32-mad", "achieved": alg_muls / (acc_ms * 1e-3),
                       "peak": modmul_peak, "unit": "Fq-mul/s", "frac": alg_muls / (acc_ms * 1e-3) / modmul_peak,
                       "imad_wide_peak_per_s": imad_peak, "mads_per_mul": 136,
                       "peak_source": "in-repo probe: register-resident Montgomery multiplications (zkb_peak_probe kind 1)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        res = cpu_prove_sample(1, 0)
        # the checker doing its job: the GPU proves the same sample and must produce the same bytes
        h2 = ctx.r1cs_load(res["r1cs"].num_constraints, res["r1cs"].num_instance, res["r1cs"].num_witness, res["r1cs"].matrices())
        p2 = ctx.pk_load(res["pk"])
        gpu_proof = ctx.prove(p2, h2, res["z"], 1234567, 7654321)
        if gpu_proof != res["proof"]:
            raise SystemExit("PARITY FAILURE: GPU proof differs from the CPU oracle on the sample circuit")
        cpu_value = res["n_cons"] / res["times"][0]
        cpu_baseline = {"value": cpu_value, "unit": "constraints/s", "cores": res["threads"], "kind": "port",
                        "sample": f"1 proof of a 2^{SAMPLE_LOG_N}-2 constraint synthetic circuit ({res['times'][0]:.1f} s, ark-equivalent C port); "
                                  "GPU proof of the same sample is byte-identical",
                        "stage_s": res["stage_s"]}

    if rank == 0:
        line = {
            "metric": "groth16_constraints_per_sec", "value": value, "unit": "constraints/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u256-montgomery", "data": "synthetic",
            "config": {"workload": f"synthetic-r1cs-2^{args.log_n}-bn128-groth16", "constraints": n_cons, "domain": 1 << args.log_n,
                       "variables": r1cs.num_variables, "witness": args.witness, "curve": "bn128",
                       "parallelism": f"msm-index-shard x{world} (work-balanced cuts), witness_map " + ("chains shared over NVLink" if share_wm else "replicated"),
                       "l2": f"inputs larger than L2: resident proving key {pk_bytes / 1e6:.0f} MB + sort buffers, no flush needed",
                       "timed_region": "z resident in HBM -> proof bytes on host",
                       "timing": "K proofs bracketed by barrier + synchronize; every proof ends in a stream synchronize inside the library, so "
                                 "the host clock equals the device time of the critical path (max over ranks); per-stage CUDA events on the "
                                 "launching streams are in stages_ms"},
            "e2e": {"value": e2e_value, "unit": "constraints/s", "h2d_bytes_per_step": int(z.nbytes + 64),
                    "d2h_bytes_per_step": int(256 + 4 * 2 * 72 * 128 + 2 * 72 * 256), "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_modmul": roofline_mm,
            "cpu_baseline": cpu_baseline,
            "stages_ms": stage_res,
            "prep_s": round(prep_s, 1),
        }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
