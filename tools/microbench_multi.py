#!/usr/bin/env python3
"""BASELINE.json config 5 at N GPUs: G1 MSM sharded by index range over the ranks (+ one all_gather of an affine point per rank,
host add), NTT replicated per GPU.  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/microbench_multi.py 20 22

Times are wall clock around the call, barrier + synchronize on both sides, max over ranks; points are the h_query of a GPU-made
proving key (distinct non-infinity points).  Prints one JSON line per size on rank 0.  (The same path is covered on gloo by tests/test_multi_gpu_gloo.py::test_sharded_msm_gloo.)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from zokrates_b200 import distributed, synthetic
    from zokrates_b200._lib import Context, Library
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", local)} if world > 1 else {}))
    dev = torch.device("cuda", local)
    sizes = [int(a) for a in sys.argv[1:]] or [18, 20]
    ctx = Context(0, local, Library())
    lg_max = min(max(sizes), int(os.environ.get("ZKB_POINTS_MAX_LOG", "24")))   # larger sizes tile this point set (timing only)
    r1cs, z = synthetic.make_layered(ctx, "bn128", (1 << lg_max) - 2)
    h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
    pk = ctx.setup(h, [11, 22, 33, 44, 5555, 3, 7])
    ctx.r1cs_free(h)
    m = r1cs.num_variables
    off = 64 + 3 * 128 + 8 + 2 * 64 + 2 * 64 + (8 + m * 64) * 2 + 8 + m * 128 + 8   # start of h_query (ark layout, ni = 2)
    hq = np.frombuffer(pk, dtype=np.uint8)[off:off + ((1 << lg_max) - 1) * 64]
    rs = np.random.RandomState(0x5EED0005 & 0x7FFFFFFF)

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for lg in sizes:
        n = (1 << lg) - 1
        reps = -(-n // ((1 << lg_max) - 1))
        pts = (np.tile(hq, reps)[:n * 64] if reps > 1 else hq[:n * 64]).tobytes()
        sc = rs.randint(0, 1 << 62, size=(n, 4)).astype(np.uint64)
        sc[:, 3] &= np.uint64((1 << 60) - 1)
        best = None
        for _ in range(4):
            sync()
            t = time.perf_counter()
            res = distributed.msm_g1_sharded(ctx, pts, sc, device=dev)
            sync()
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        tt = torch.tensor([best], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        x = rs.randint(0, 1 << 62, size=(1 << lg, 4)).astype(np.uint64)
        x[:, 3] &= np.uint64((1 << 60) - 1)
        ctx.ntt(x)
        ntt_ms = ctx.timings()["ntt"]
        if rank == 0:
            ms = float(tt.item()) * 1e3
            print(json.dumps({"log_n": lg, "n_gpus": world, "points_tiled": reps > 1, "msm_g1_ms_incl_h2d": ms, "msm_fq_mul_per_s": n * 16 * 10 / (ms * 1e-3),
                              "ntt_ms_per_gpu": ntt_ms, "result_sha": __import__("hashlib").sha256(res).hexdigest()[:16]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
