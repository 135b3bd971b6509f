"""GPU tier (-m gpu): parity AT THE SIZES bench.py TIMES — the 2^20-constraint proofs (BASELINE configs 3 / 4 family) and
the window-table MSM path — against two independent CPU answers, through the C ABI:

  * `zko_groth16_prove`     : the ark-equivalent CPU prover (oracle/zkoracle.c: ark's MSM / FFT schedule on the key bytes)
  * `zko_trapdoor_expected` : the proof predicted from the setup trapdoor with Fr arithmetic only (closed-form Lagrange
                              coefficients, three generator multiplications — no NTT, no MSM, no proving key)

Proof bytes are canonical (affine points), so equality is bit-exactness.  Nothing here reads /root/reference."""
import numpy as np
import pytest

from oracle.ff import BLS12_381, BN254
from zokrates_b200 import synthetic
from zokrates_b200._lib import (OPT_TABLE_C, OPT_TABLE_MIN_LOG, OPT_TABLES, OPT_Z_MODE, Context, ZkbError)

pytestmark = pytest.mark.gpu
TD = [3, 5, 7, 11, 1234567, 17, 19]
R, S = 1234567, 7654321


class Circuit:
    def __init__(self, lib, cid, c, log_n, dist):
        self.cid, self.c = cid, c
        self.ctx = Context(cid, 0, lib)
        self.r1, self.z = synthetic.make_layered(self.ctx, c.name, (1 << log_n) - 2, distribution=dist)
        self.h = self.ctx.r1cs_load(self.r1.num_constraints, self.r1.num_instance, self.r1.num_witness, self.r1.matrices())
        self.pk = self.ctx.setup(self.h, TD)

    def close(self):
        self.ctx.close()


@pytest.fixture(scope="module", params=[(0, BN254, "uniform"), (0, BN254, "bits"), (1, BLS12_381, "uniform")],
                ids=lambda p: f"{p[1].name}-{p[2]}")
def full(request, gpu_lib):
    cid, c, dist = request.param
    circ = Circuit(gpu_lib, cid, c, 20, dist)
    yield circ
    circ.close()


def test_full_size_proof_vs_oracle_and_trapdoor(full, oracle_c):
    """2^20 - 2 constraints (what bench.py proves): GPU proof == ark-equivalent CPU proof == trapdoor prediction, in both
    MSM modes of the assignment MSMs (shared-bucket window tables / per-window buckets) and with the tables switched off."""
    ctx, c = full.ctx, full.c
    expected = oracle_c.trapdoor_expected(full.cid, full.r1, TD, full.z, R, S, c.fq_bytes)
    assert ctx.r1cs_check(full.h, full.z) is None
    pkh = ctx.pk_load(full.pk)
    info = ctx.pk_table_info(pkh)
    assert info["z_tables"] == "built" and info["h_table"] == "built" and info["c_z"] >= 16, info
    proof = ctx.prove(pkh, full.h, full.z, R, S)
    assert proof == expected, "GPU proof differs from the trapdoor prediction"
    cpu, _ = oracle_c.prove(full.cid, full.pk, full.r1, full.z, R, S, c.fq_bytes)
    assert cpu == expected, "CPU oracle differs from the trapdoor prediction"
    for mode in (1, 2):                       # the sampling switch assignment_is_sparse picks one of these per witness
        ctx.set_option(OPT_Z_MODE, mode)
        assert ctx.prove(pkh, full.h, full.z, R, S) == expected, f"z mode {mode}"
    ctx.set_option(OPT_Z_MODE, 0)
    ctx.pk_free(pkh)
    ctx.set_option(OPT_TABLES, 0)
    pkh = ctx.pk_load(full.pk)
    assert ctx.pk_table_info(pkh)["table_bytes"] == 0
    assert ctx.prove(pkh, full.h, full.z, R, S) == expected, "no-table path"
    ctx.pk_free(pkh)
    ctx.set_option(OPT_TABLES, 1)


def test_full_size_sharding_vs_trapdoor(full, oracle_c):
    """1-, 2- and 8-way index sharding of the 2^20 proof (what the multi-GPU bench does, here on one GPU) all give the
    trapdoor-predicted bytes; a different r gives a different proof (the blinding is applied)."""
    ctx, c = full.ctx, full.c
    if full.cid != 0:
        pytest.skip("sharding invariance is curve independent: BN254 only")
    expected = oracle_c.trapdoor_expected(full.cid, full.r1, TD, full.z, R, S, c.fq_bytes)
    pkh = ctx.pk_load(full.pk)
    for world in (2, 8):
        parts = []
        for rank in range(world):
            ph = ctx.pk_load(full.pk, rank, world)
            parts.append(ctx.prove_partial(ph, full.h, full.z))
            ctx.pk_free(ph)
        assert ctx.finalize(pkh, np.concatenate(parts), world, R, S) == expected
    assert ctx.prove(pkh, full.h, full.z, R + 1, S) != expected
    ctx.pk_free(pkh)


def test_full_size_pipelined_proofs(full, oracle_c):
    """zkb_groth16_prove_submit / _collect with TWO 2^20 proofs in flight (what bench.py times): host-memory assignment and
    resident assignment, different blinding scalars per proof, collected in both orders — every proof equals the trapdoor
    prediction for its own (r, s)."""
    ctx, c = full.ctx, full.c
    exp = {rs: oracle_c.trapdoor_expected(full.cid, full.r1, TD, full.z, rs[0], rs[1], c.fq_bytes) for rs in ((R, S), (5, 7), (11, 13))}
    pkh = ctx.pk_load(full.pk)
    ctx.set_assignment(full.h, full.z)
    t1 = ctx.prove_submit(pkh, full.h, full.z, R, S)
    t2 = ctx.prove_submit(pkh, full.h, None, 5, 7)
    assert ctx.prove_collect(t1) == exp[(R, S)]
    t3 = ctx.prove_submit(pkh, full.h, full.z, 11, 13)
    assert ctx.prove_collect(t3) == exp[(11, 13)]          # out of order
    assert ctx.prove_collect(t2) == exp[(5, 7)]
    # a stream of proofs, depth 2
    pending, got = [], []
    for i in range(6):
        pending.append(ctx.prove_submit(pkh, full.h, None if i % 2 else full.z, R, S))
        if len(pending) == 2:
            got.append(ctx.prove_collect(pending.pop(0)))
    got += [ctx.prove_collect(t) for t in pending]
    assert got == [exp[(R, S)]] * 6
    ctx.pk_free(pkh)


@pytest.mark.parametrize("log_n", [15, 17])
@pytest.mark.parametrize("dist", ["uniform", "bits"])
def test_table_modes_vs_oracle(gpu_lib, oracle_c, log_n, dist):
    """The window-table path at mid sizes against the ark-equivalent CPU prover: tables off, cost-model window, forced
    narrow / wide windows (c = 13 -> W = 20 is refused: W <= 16; c = 16, 21), both assignment-MSM modes."""
    circ = Circuit(gpu_lib, 0, BN254, log_n, dist)
    ctx, c = circ.ctx, circ.c
    ref, _ = oracle_c.prove(0, circ.pk, circ.r1, circ.z, R, S, c.fq_bytes)
    assert ref == oracle_c.trapdoor_expected(0, circ.r1, TD, circ.z, R, S, c.fq_bytes)
    seen = set()
    for tables, tc in ((0, 0), (1, 0), (1, 16), (1, 21), (1, 13)):
        ctx.set_option(OPT_TABLES, tables)
        ctx.set_option(OPT_TABLE_C, tc)
        pkh = ctx.pk_load(circ.pk)
        info = ctx.pk_table_info(pkh)
        seen.add((info["z_tables"], info["c_z"]))
        if tables and tc in (16, 21):
            assert info["c_z"] == tc and info["c_h"] == tc and info["z_tables"] == "built"
        if tc == 13:
            assert info["z_tables"] == "no-window" and info["table_bytes"] == 0
        for mode in (0, 1, 2):
            ctx.set_option(OPT_Z_MODE, mode)
            assert ctx.prove(pkh, circ.h, circ.z, R, S) == ref, (tables, tc, mode)
        ctx.pk_free(pkh)
    assert len(seen) >= 4
    circ.close()


def test_table_fallback_when_hbm_is_short(gpu_lib, oracle_c):
    """The bytes-for-multiplications trade must degrade visibly, never silently: with HBM nearly full the z tables are
    reported 'no-memory' (the h table, 5x smaller, still fits), the proof is unchanged, and ZKB_OPT_TABLES = 2 turns the
    same situation into ZKB_E_OOM."""
    import torch
    from zokrates_b200._lib import OPT_PK_CACHE
    circ = Circuit(gpu_lib, 0, BN254, 18, "uniform")
    ctx, c = circ.ctx, circ.c
    ctx.set_option(OPT_PK_CACHE, 0)            # every load must build (or fail to build) its own tables here
    expected = oracle_c.trapdoor_expected(0, circ.r1, TD, circ.z, R, S, c.fq_bytes)
    pkh = ctx.pk_load(circ.pk)
    full_info = ctx.pk_table_info(pkh)
    assert full_info["z_tables"] == "built"
    assert ctx.prove(pkh, circ.h, circ.z, R, S) == expected
    ctx.pk_free(pkh)
    need_z = full_info["table_bytes"] * 5 // 6          # a/b1/l (3 x 64 B) + b2 (128 B) of 6 x 64 B per (point, window)
    torch.cuda.synchronize()
    free_b, _ = torch.cuda.mem_get_info(0)
    # leave: the reserve the library keeps (sort plans + 1 GiB) + half of what the z tables need
    keep = (1 << 30) + 16 * 17 * 2 * (1 << 18) + 6 * 32 * (1 << 18) + need_z // 2 + (256 << 20)
    hog = torch.empty(free_b - keep, dtype=torch.uint8, device="cuda:0")
    try:
        pkh = ctx.pk_load(circ.pk)
        info = ctx.pk_table_info(pkh)
        assert info["z_tables"] == "no-memory" and info["c_z"] == 0, info
        assert info["h_table"] == "built", info
        assert ctx.prove(pkh, circ.h, circ.z, R, S) == expected
        ctx.pk_free(pkh)
        ctx.set_option(OPT_TABLES, 2)
        with pytest.raises(ZkbError) as e:
            ctx.pk_load(circ.pk)
        assert e.value.code == 4 and "window tables" in str(e.value)
    finally:
        del hog
        torch.cuda.empty_cache()
        ctx.set_option(OPT_TABLES, 1)
        ctx.set_option(OPT_PK_CACHE, 1)
    circ.close()


def _pk_sections(c, pk):
    n = c.fq_bytes
    off = 2 * n + 3 * 4 * n
    cnt = int.from_bytes(pk[off:off + 8], "little"); off += 8 + cnt * 2 * n + 2 * 2 * n
    m = int.from_bytes(pk[off:off + 8], "little"); off += 8
    a = pk[off:off + m * 2 * n]; off += m * 2 * n + 8
    off += m * 2 * n + 8
    b2 = pk[off:off + m * 4 * n]
    return a, b2, m


@pytest.mark.parametrize("cid,c", [(0, BN254), (1, BLS12_381)], ids=lambda p: getattr(p, "name", ""))
def test_standalone_msm_vs_c_oracle_large(gpu_lib, oracle_c, cid, c):
    """zkb_msm_g1 at 2^16 and 2^18 pairs, zkb_msm_g2 at 2^16 (BASELINE config 5 entry points), uniform and bit-heavy
    scalars, against ark's VariableBaseMSM schedule on the CPU."""
    ctx = Context(cid, 0, gpu_lib)
    r1, _ = synthetic.make_layered(ctx, c.name, (1 << 18) - 8)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    pk = ctx.setup(h, TD)
    ctx.r1cs_free(h)
    a, b2, m = _pk_sections(c, pk)
    assert m == 1 << 18
    rs = np.random.RandomState(31)
    sc = rs.randint(0, 1 << 62, size=(m, 4)).astype(np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)
    bits = sc.copy()
    small = rs.rand(m) < 0.9
    bits[small] = 0
    bits[small, 0] = rs.randint(0, 2, size=int(small.sum())).astype(np.uint64)
    g1b, g2b = 2 * c.fq_bytes, 4 * c.fq_bytes
    for scalars in (sc, bits):
        for n in (1 << 16, 1 << 18):
            assert ctx.msm(1, a[:n * g1b], scalars[:n]) == oracle_c.msm(cid, 1, a[:n * g1b], scalars[:n], c.fq_bytes), n
        n = 1 << 16
        assert ctx.msm(2, b2[:n * g2b], scalars[:n]) == oracle_c.msm(cid, 2, b2[:n * g2b], scalars[:n], c.fq_bytes)
    ctx.close()


def test_pk_cache_by_content(gpu_lib, oracle_c):
    """ZKB_OPT_PK_CACHE: the trait-shaped pk_load / prove / pk_free per proof builds the window tables once — a second load of the
    same bytes returns a handle onto the resident key (also after the last handle was released), other bytes or other table
    options do not hit; proofs are unchanged."""
    import time
    circ = Circuit(gpu_lib, 0, BN254, 16, "uniform")
    ctx, c = circ.ctx, circ.c
    expected = oracle_c.trapdoor_expected(0, circ.r1, TD, circ.z, R, S, c.fq_bytes)
    t = time.perf_counter(); h1 = ctx.pk_load(circ.pk); cold = time.perf_counter() - t
    assert "pk_cache_hit" not in ctx.timings()
    h2 = ctx.pk_load(circ.pk)
    assert h2 != h1 and "pk_cache_hit" in ctx.timings()
    assert ctx.prove(h2, circ.h, circ.z, R, S) == expected
    ctx.pk_free(h1); ctx.pk_free(h2)
    with pytest.raises(ZkbError):
        ctx.pk_table_info(h2)
    t = time.perf_counter(); h3 = ctx.pk_load(circ.pk); warm = time.perf_counter() - t
    assert "pk_cache_hit" in ctx.timings() and warm < cold
    assert ctx.prove(h3, circ.h, circ.z, R, S) == expected
    ctx.pk_free(h3)
    other = bytearray(circ.pk); other[100] ^= 1                      # different bytes (a coordinate of beta_g2): a miss
    h4 = ctx.pk_load(bytes(other))
    assert "pk_cache_hit" not in ctx.timings()
    ctx.pk_free(h4)
    ctx.set_option(OPT_TABLES, 0)
    h5 = ctx.pk_load(circ.pk)                                          # other table options: a miss, no tables
    assert "pk_cache_hit" not in ctx.timings() and ctx.pk_table_info(h5)["z_tables"] == "disabled"
    assert ctx.prove(h5, circ.h, circ.z, R, S) == expected
    ctx.pk_free(h5)
    ctx.set_option(OPT_TABLES, 1)
    circ.close()
