"""N > 1 path on CPU: world_size-2/3 gloo process groups drive the sharded prover (host-emulation engine);
the gathered proof must equal the single-rank proof byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, emu_path, out_path, curve="bn128"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zokrates_b200 import backend, distributed, synthetic
        from zokrates_b200._lib import Library
        lib = Library(emu_path)
        r1cs, z = synthetic.make(curve, 60, seed=3)
        ctx0 = backend.context(curve, 0, lib)
        h = ctx0.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
        pk = ctx0.setup(h, [3, 5, 7, 11, 13, 17, 19])
        sess = backend.ProverSession(curve, r1cs, pk, 0, rank, world, lib=lib)
        proof = distributed.prove_sharded(sess, z, 111, 222)
        if rank == 0:
            single = backend.ProverSession(curve, r1cs, pk, 0, 0, 1, lib=lib).prove_raw(z, 111, 222)
            np.save(out_path, np.array([proof == single, len(proof)]))
        else:
            assert proof is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_prove_gloo(world, emu_lib, tmp_path):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), emu_lib.path, out), nprocs=world, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n == 256


def _worker_chains(rank, world, port, emu_path, out_path):
    """begin/end with the chain split: every rank computes only its chains, the others arrive by broadcast.  Without the
    exchange the result must differ on the ranks that do not own chain a (so the test cannot pass on replicated work)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from zokrates_b200 import backend, distributed, synthetic
        from zokrates_b200._lib import Library, ZkbError
        lib = Library(emu_path)
        r1cs, z = synthetic.make("bn128", 1500, seed=5)          # domain 2^11: the tiled NTT path
        ctx = backend.context("bn128", 0, lib)
        h = ctx.r1cs_load(r1cs.num_constraints, r1cs.num_instance, r1cs.num_witness, r1cs.matrices())
        pk = ctx.setup(h, [3, 5, 7, 11, 13, 17, 19])
        pkh = ctx.pk_load(pk, rank, world)
        ref = ctx.prove_partial(pkh, h, z)
        mask = distributed.wm_chain_mask(rank, world)
        assert mask == ({0: 1, 1: 2, 2: 4}.get(rank, 0))
        got = distributed.prove_partial_shared_wm(ctx, pkh, h, z)
        ok = bool(np.array_equal(ref, got))
        ctx.prove_begin(pkh, h, z, mask)                           # no exchange this time
        try:
            ctx.prove_begin(pkh, h, z, mask)
            ok = False
        except ZkbError as e:
            ok = ok and e.code == 1                                # "a proof is already open"
        stale = ctx.prove_end(pkh, h)
        # the finish step consumed buffer a: only its owner (rank 0) recomputed it, everyone else now has a wrong h
        ok = ok and (bool(np.array_equal(stale, ref)) == (rank == 0))
        try:
            ctx.prove_end(pkh, h)
            ok = False
        except ZkbError as e:
            ok = ok and e.code == 1                                # end without begin
        np.save(out_path + f".{rank}.npy", np.array([ok]))
    finally:
        dist.destroy_process_group()


def test_sharded_prove_gloo_bls12_381(emu_lib, tmp_path):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(3, _free_port(), emu_lib.path, out, "bls12_381"), nprocs=3, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n == 384


@pytest.mark.parametrize("world", [3, 4])
def test_shared_witness_map_chains_gloo(world, emu_lib, tmp_path):
    out = str(tmp_path / "res")
    mp.spawn(_worker_chains, args=(world, _free_port(), emu_lib.path, out), nprocs=world, join=True)
    for rank in range(world):
        assert np.load(out + f".{rank}.npy")[0] == 1, f"rank {rank}"


def _worker_msm(rank, world, port, emu_path, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import random
        from oracle import ark
        from oracle.ff import BN254, g1_group
        from zokrates_b200 import backend, distributed
        from zokrates_b200._lib import Library, fr_array
        lib = Library(emu_path)
        ctx = backend.context("bn128", 0, lib)
        rnd = random.Random(9)
        G1 = g1_group(BN254)
        pts = [G1.mul(BN254.g1, rnd.randrange(1, BN254.r)) for _ in range(11)]
        pts[4] = None
        sc = [rnd.choice([0, 1, BN254.r - 1, rnd.randrange(BN254.r)]) for _ in range(11)]
        got = distributed.msm_g1_sharded(ctx, b"".join(ark.ser_g1(BN254, p) for p in pts), fr_array(sc))
        ok = got == ark.ser_g1(BN254, G1.msm_naive(pts, sc))
        zero = distributed.msm_g1_sharded(ctx, b"".join(ark.ser_g1(BN254, p) for p in pts), fr_array([0] * 11))
        ok = ok and zero == ark.ser_g1(BN254, None)
        np.save(out_path + f".{rank}.npy", np.array([ok]))
    finally:
        dist.destroy_process_group()


def test_sharded_msm_gloo(emu_lib, tmp_path):
    out = str(tmp_path / "msm")
    mp.spawn(_worker_msm, args=(3, _free_port(), emu_lib.path, out), nprocs=3, join=True)
    for rank in range(3):
        assert np.load(out + f".{rank}.npy")[0] == 1, f"rank {rank}"
