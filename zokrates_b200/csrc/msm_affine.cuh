// Batch-affine rounds in front of the bucket accumulation.  EXPERIMENTAL, OFF BY DEFAULT (ZKB_OPT_BATCH_AFFINE = 0): measured on
// B200 the rounds are 3.3x slower than the direct XYZZ path (2^20 points: 5.2 + 2.4 + 1.3 + 0.3 ms against 2.5 ms) — the single
// Fermat inversion per block (~380 dependent multiplications on one thread, ~0.1 ms) stalls 127 threads, and with four blocks
// per SM the stalls do not interleave; a three-kernel split (products / parallel inversions / additions) would remove the stall
// but at 7.4 multiplications per addition with two passes over the points it lands at parity with the direct path at best
// (profiles/r02_batch_affine.md).  Kept because it is correct, tested (tests/test_batch_affine.py) and the starting point for that
// split.
//
// The mixed addition into an XYZZ accumulator costs 10 field multiplications (8M + 2S; 28 base-field multiplications in
// G2).  Adding two AFFINE points costs 1 inversion + 2M + 1S, and Montgomery's trick turns k inversions into one inversion
// and 3(k - 1) multiplications: 6 multiplications per addition (17 instead of 28 in G2) once k is in the thousands.  The
// accumulation kernels sit at 0.81-0.88 of the integer-multiply pipe that binds them (profiles/r02_ncu_accum1_g2.md), so
// fewer multiplications is the lever left.
//
// The sorted (bucket, point) list makes independent additions available at no cost: inside a bucket, entries 2q and 2q + 1
// can be added pairwise.  One ROUND halves every bucket: segment b of length L becomes ceil(L / 2) affine points (an odd
// last entry is copied), offsets' = scan(ceil(L / 2)).  After R rounds (R = 3: 7/8 of the additions) the shortened list
// goes through the unchanged XYZZ chunk accumulation (msm_accum1_body), which also absorbs any skew (a bucket of any length
// is just a long segment), followed by the same segmented partial levels and bit-sum reduction.  Results are sums in the
// group: bit-identical to the direct path.
//
// One block = 128 threads x BA_M consecutive outputs each.  Forward pass: denominators d (x2 - x1, or 2 y for a doubling, 1
// for copies / cancellations), running product per thread, prefixes parked in shared memory.  Block product tree in shared
// memory (127 multiplications up, 254 down), ONE Fermat inversion per block by thread 0.  Backward pass: 2 multiplications
// recover each 1 / d, 1M + 1S + 1M finish the addition.
#pragma once
#include <vector>
#include "msm.cuh"

namespace zkb {

static constexpr uint32_t BA_M = 8;            // outputs per thread
static constexpr uint32_t BA_BLOCK = 128;
static constexpr uint32_t BA_TILE = BA_M * BA_BLOCK;

// ceil(L / 2) per bucket (input of the scan that gives the next round's offsets)
ZKB_HDN inline void ba_halve_counts_body(uint32_t NB, const uint32_t* off_in, uint32_t* cnt_out, uint32_t b) {
  if (b >= NB) return;
  const uint32_t L = off_in[b + 1] - off_in[b];
  cnt_out[b] = (L + 1) >> 1;
}

// the bucket that holds output j: largest b with off_out[b] <= j (off_out non-decreasing, j < off_out[NB])
ZKB_HD uint32_t ba_find_bucket(const uint32_t* off_out, uint32_t NB, uint32_t j) {
  uint32_t lo = 0, hi = NB;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (off_out[mid] <= j) lo = mid; else hi = mid;
  }
  while (off_out[lo + 1] <= j) lo++;   // empty buckets sharing the offset
  return lo;
}

enum BaMode : uint32_t { BA_ADD = 0, BA_DBL = 1, BA_COPY1 = 2, BA_COPY2 = 3, BA_INF = 4 };

// load input point `pos` of a round: round 0 reads the point table through the sorted list (sign bit = negate), later rounds
// read the previous round's output directly
template <class F>
ZKB_HD Affine<F> ba_load(const uint32_t* sorted, const Affine<F>* pts, uint32_t pos) {
  if (!sorted) return pts[pos];
  const uint32_t e = sorted[pos];
  Affine<F> p = pts[e & ~MSM_NEG];
  if (e & MSM_NEG) p.y = F::neg(p.y);
  return p;
}

// mode and denominator of one pair (p2 meaningful only when has2)
template <class F>
ZKB_HD uint32_t ba_mode(const Affine<F>& p1, const Affine<F>& p2, bool has2, F& den) {
  den = F::one();
  if (!has2 || p2.is_inf()) return BA_COPY1;
  if (p1.is_inf()) return BA_COPY2;
  F dx = F::sub(p2.x, p1.x);
  if (!dx.is_zero()) { den = dx; return BA_ADD; }
  if (p1.y == p2.y && !p1.y.is_zero()) { den = F::dbl(p1.y); return BA_DBL; }
  return BA_INF;                         // P + (-P)
}

// finish one pair given 1 / den
template <class F>
ZKB_HD Affine<F> ba_finish(uint32_t mode, const Affine<F>& p1, const Affine<F>& p2, const F& dinv) {
  if (mode == BA_COPY1) return p1;
  if (mode == BA_COPY2) return p2;
  if (mode == BA_INF) return Affine<F>::inf();
  F num;
  if (mode == BA_DBL) { F xx = F::sqr(p1.x); num = F::add(F::dbl(xx), xx); }
  else num = F::sub(p2.y, p1.y);
  const F lam = F::mul(num, dinv);
  const F x3 = F::sub(F::sub(F::sqr(lam), p1.x), mode == BA_DBL ? p1.x : p2.x);
  const F y3 = F::sub(F::mul(lam, F::sub(p1.x, x3)), p1.y);
  return Affine<F>{x3, y3};
}

// reference form of one output (own inversion): the host emulation runs this, the device kernel below batches the inversions
template <class F>
ZKB_HDN inline void ba_round_ref_body(uint32_t NB, const uint32_t* off_in, const uint32_t* off_out, const uint32_t* sorted,
                                      const Affine<F>* pts_in, Affine<F>* pts_out, uint32_t j) {
  if (j >= off_out[NB]) return;
  const uint32_t b = ba_find_bucket(off_out, NB, j);
  const uint32_t pin = off_in[b] + 2 * (j - off_out[b]);
  const bool has2 = pin + 1 < off_in[b + 1];
  const Affine<F> p1 = ba_load<F>(sorted, pts_in, pin);
  Affine<F> p2 = p1;
  if (has2) p2 = ba_load<F>(sorted, pts_in, pin + 1);
  F den;
  const uint32_t mode = ba_mode<F>(p1, p2, has2, den);
  pts_out[j] = ba_finish<F>(mode, p1, p2, mode <= BA_DBL ? F::inv(den) : den);
}

// ---- the batched block: shared by the device kernel and the host emulation ---------------------------------------------
// Block-shared vector of field elements.  Device: 16-byte planes (plane q of element i at [q * count + i], conflict-free for
// unit-stride i); host emulation: a plain array.
#if !defined(ZKB_EMU)
template <class F>
struct BaVec {
  static constexpr int Q = sizeof(F) / 16;
  static_assert(sizeof(F) % 16 == 0, "16-byte planes");
  uint4* base;
  uint32_t count;
  __device__ __forceinline__ void st(uint32_t i, const F& v) const {
    const uint4* s = (const uint4*)&v;
#pragma unroll
    for (int q = 0; q < Q; q++) base[(size_t)q * count + i] = s[q];
  }
  __device__ __forceinline__ F ld(uint32_t i) const {
    F v;
    uint4* d = (uint4*)&v;
#pragma unroll
    for (int q = 0; q < Q; q++) d[q] = base[(size_t)q * count + i];
    return v;
  }
};
#else
template <class F>
struct BaVec {
  F* base;
  uint32_t count;
  void st(uint32_t i, const F& v) const { base[i] = v; }
  F ld(uint32_t i) const { return base[i]; }
};
#endif

template <class F>
struct BaShared {
  BaVec<F> pref;    // BA_TILE prefixes, index k * BA_BLOCK + tid
  BaVec<F> node;    // heap-ordered product tree over the thread products: nodes 1 .. 2 BA_BLOCK - 1, leaves at BA_BLOCK + tid
  BaVec<F> ninv;    // inverses of the same nodes
};
template <class F>
struct BaThread {   // what a thread keeps between the passes (registers on the device)
  uint32_t pin[BA_M];   // input position of the first point | has2 << 31; ~0: no output
  F run;
};

// forward pass of thread `tid` of block `blk`: locate the pairs, multiply the denominators up, park the prefixes
template <class F>
ZKB_HD void ba_forward(BaThread<F>& th, const BaShared<F>& sh, uint32_t NB, const uint32_t* off_in, const uint32_t* off_out,
                       const uint32_t* sorted, const Affine<F>* pts_in, uint32_t blk, uint32_t tid) {
  const uint32_t M = off_out[NB];
  const uint32_t j0 = (blk * BA_BLOCK + tid) * BA_M;
  F run = F::one();
  uint32_t b = 0, ob = 0, ob1 = 0, ib = 0, ib1 = 0;
  if (j0 < M) { b = ba_find_bucket(off_out, NB, j0); ob = off_out[b]; ob1 = off_out[b + 1]; ib = off_in[b]; ib1 = off_in[b + 1]; }
#pragma unroll
  for (uint32_t k = 0; k < BA_M; k++) {
    const uint32_t j = j0 + k;
    sh.pref.st(k * BA_BLOCK + tid, run);
    if (j >= M) { th.pin[k] = 0xFFFFFFFFu; continue; }
    while (ob1 <= j) { b++; ob = ob1; ob1 = off_out[b + 1]; ib = ib1; ib1 = off_in[b + 1]; }
    const uint32_t p = ib + 2 * (j - ob);
    const bool has2 = p + 1 < ib1;
    th.pin[k] = p | (has2 ? 0x80000000u : 0u);
    const Affine<F> p1 = ba_load<F>(sorted, pts_in, p);
    Affine<F> p2 = p1;
    if (has2) p2 = ba_load<F>(sorted, pts_in, p + 1);
    F den;
    const uint32_t mode = ba_mode<F>(p1, p2, has2, den);
    if (mode <= BA_DBL) run = F::mul(run, den);
  }
  th.run = run;
  sh.node.st(BA_BLOCK + tid, run);
}
// product tree, one level: nodes s .. 2s - 1 (threads tid < s)
template <class F>
ZKB_HD void ba_tree_up(const BaShared<F>& sh, uint32_t s, uint32_t tid) {
  if (tid >= s) return;
  const uint32_t i = s + tid;
  sh.node.st(i, F::mul(sh.node.ld(2 * i), sh.node.ld(2 * i + 1)));
}
// the one inversion of the block (every factor is a non-zero denominator or 1, so the root is never zero)
template <class F>
ZKB_HD void ba_tree_root(const BaShared<F>& sh, uint32_t tid) {
  if (tid == 0) sh.ninv.st(1, F::inv(sh.node.ld(1)));
}
// inverse tree, one level: children of nodes s .. 2s - 1:  1/L = (1/LR) R,  1/R = (1/LR) L
template <class F>
ZKB_HD void ba_tree_down(const BaShared<F>& sh, uint32_t s, uint32_t tid) {
  if (tid >= s) return;
  const uint32_t i = s + tid;
  const F pi = sh.ninv.ld(i);
  sh.ninv.st(2 * i, F::mul(pi, sh.node.ld(2 * i + 1)));
  sh.ninv.st(2 * i + 1, F::mul(pi, sh.node.ld(2 * i)));
}
// backward pass: 2 multiplications recover each 1 / den, then the addition
template <class F>
ZKB_HD void ba_backward(const BaThread<F>& th, const BaShared<F>& sh, const uint32_t* sorted, const Affine<F>* pts_in, Affine<F>* pts_out,
                        uint32_t blk, uint32_t tid) {
  const uint32_t j0 = (blk * BA_BLOCK + tid) * BA_M;
  F inv = sh.ninv.ld(BA_BLOCK + tid);                      // 1 / (product of this thread's denominators)
#pragma unroll
  for (uint32_t kk = 0; kk < BA_M; kk++) {
    const uint32_t k = BA_M - 1 - kk;
    if (th.pin[k] == 0xFFFFFFFFu) continue;
    const uint32_t p = th.pin[k] & 0x7FFFFFFFu;
    const bool has2 = (th.pin[k] >> 31) != 0;
    const Affine<F> p1 = ba_load<F>(sorted, pts_in, p);
    Affine<F> p2 = p1;
    if (has2) p2 = ba_load<F>(sorted, pts_in, p + 1);
    F den;
    const uint32_t mode = ba_mode<F>(p1, p2, has2, den);
    F dinv = den;
    if (mode <= BA_DBL) {
      dinv = F::mul(inv, sh.pref.ld(k * BA_BLOCK + tid));  // 1 / den = (1 / prefix_{k+1}) * prefix_k
      inv = F::mul(inv, den);
    }
    pts_out[j0 + k] = ba_finish<F>(mode, p1, p2, dinv);
  }
}

template <class F>
constexpr size_t ba_smem_bytes() { return (size_t)(BA_TILE + 2 * 2 * BA_BLOCK) * sizeof(F); }   // prefixes + product tree + inverse tree

#if !defined(ZKB_EMU)
template <class F, int MINB>
__global__ void __launch_bounds__(BA_BLOCK, MINB) zkb_batch_affine(uint32_t NB, const uint32_t* __restrict__ off_in,
                                                                    const uint32_t* __restrict__ off_out, const uint32_t* __restrict__ sorted,
                                                                    const Affine<F>* __restrict__ pts_in, Affine<F>* __restrict__ pts_out) {
  extern __shared__ uint4 ba_smem[];
  constexpr int Q = BaVec<F>::Q;
  BaShared<F> sh;
  sh.pref = BaVec<F>{ba_smem, BA_TILE};
  sh.node = BaVec<F>{ba_smem + (size_t)Q * BA_TILE, 2 * BA_BLOCK};
  sh.ninv = BaVec<F>{ba_smem + (size_t)Q * (BA_TILE + 2 * BA_BLOCK), 2 * BA_BLOCK};
  if ((uint64_t)blockIdx.x * BA_TILE >= off_out[NB]) return;       // whole block beyond the list (uniform)
  const uint32_t tid = threadIdx.x;
  BaThread<F> th;
  ba_forward<F>(th, sh, NB, off_in, off_out, sorted, pts_in, blockIdx.x, tid);
  __syncthreads();
  for (uint32_t s = BA_BLOCK >> 1; s >= 1; s >>= 1) { ba_tree_up<F>(sh, s, tid); __syncthreads(); }
  ba_tree_root<F>(sh, tid);
  __syncthreads();
  for (uint32_t s = 1; s < BA_BLOCK; s <<= 1) { ba_tree_down<F>(sh, s, tid); __syncthreads(); }
  ba_backward<F>(th, sh, sorted, pts_in, pts_out, blockIdx.x, tid);
}
#else
// host emulation of one block: the same passes, thread after thread between the barriers
template <class F>
inline void ba_block_emulate(uint32_t NB, const uint32_t* off_in, const uint32_t* off_out, const uint32_t* sorted, const Affine<F>* pts_in,
                             Affine<F>* pts_out, uint32_t blk) {
  if ((uint64_t)blk * BA_TILE >= off_out[NB]) return;
  std::vector<F> pref(BA_TILE), node(2 * BA_BLOCK), ninv(2 * BA_BLOCK);
  std::vector<BaThread<F>> th(BA_BLOCK);
  BaShared<F> sh{BaVec<F>{pref.data(), BA_TILE}, BaVec<F>{node.data(), 2 * BA_BLOCK}, BaVec<F>{ninv.data(), 2 * BA_BLOCK}};
  for (uint32_t t = 0; t < BA_BLOCK; t++) ba_forward<F>(th[t], sh, NB, off_in, off_out, sorted, pts_in, blk, t);
  for (uint32_t s = BA_BLOCK >> 1; s >= 1; s >>= 1) for (uint32_t t = 0; t < BA_BLOCK; t++) ba_tree_up<F>(sh, s, t);
  for (uint32_t t = 0; t < BA_BLOCK; t++) ba_tree_root<F>(sh, t);
  for (uint32_t s = 1; s < BA_BLOCK; s <<= 1) for (uint32_t t = 0; t < BA_BLOCK; t++) ba_tree_down<F>(sh, s, t);
  for (uint32_t t = 0; t < BA_BLOCK; t++) ba_backward<F>(th[t], sh, sorted, pts_in, pts_out, blk, t);
}
#endif

}  // namespace zkb
