// Radix-2 NTT over the scalar field and the pointwise steps of the QAP witness map.
//
// Replaces ark-poly 0.3.0 `Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place` and
// `divide_by_vanishing_poly_on_coset_in_place` as used by ark-groth16's
// `LibsnarkReduction::witness_map` (external, Cargo.lock:282,221; reached from
// /root/reference/zokrates_ark/src/groth16.rs:44).  The domain
// generator is ark's: omega = TWO_ADIC_ROOT^(2^(S - log n)), coset shift g = multiplicative generator.
//
// Schedule: decimation-in-frequency passes (natural -> bit-reversed) for the inverse transforms and
// decimation-in-time passes (bit-reversed -> natural) for the forward coset transforms, so no
// bit-reversal pass is needed between ifft and coset_fft.  Each pass does up to 3 butterfly stages
// on 8 elements held in registers (one read + one write of the vector per pass); twiddles come from
// a per-domain table of omega^k, k < n/2, built once on the device.
#pragma once
#include "fp.cuh"

namespace zkb {

ZKB_HD uint32_t bitrev32(uint32_t x, uint32_t bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
  uint32_t r = 0;
  for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

// tw[k] = w^k for k < count (w given in Montgomery form); each thread exponentiates independently.
template <class Fr>
ZKB_HDN inline void ntt_powers_body(Fr w, Fr scale, Fr* out, uint32_t count, uint32_t t) {
  if (t >= count) return;
  out[t] = Fr::mul(scale, Fr::pow_u64(w, t));
}

// One DIF pass: K stages with half-spans h0, h0/2, ..., h0 >> (K-1).   (natural -> bit-reversed)
template <class Fr, int K>
ZKB_HDN inline void ntt_dif_body(Fr* x, const Fr* tw, uint32_t log_n, uint32_t h0, uint32_t t) {
  constexpr uint32_t R = 1u << K;
  const uint32_t n = 1u << log_n;
  if (t >= (n >> K)) return;
  const uint32_t hmin = h0 >> (K - 1);
  const uint32_t off = t & (hmin - 1), blk = t / hmin;
  const size_t i0 = (size_t)blk * (2 * (size_t)h0) + off;
  Fr e[R];
#pragma unroll
  for (uint32_t m = 0; m < R; m++) e[m] = x[i0 + (size_t)m * hmin];
#pragma unroll
  for (int q = 0; q < K; q++) {
    const uint32_t hm = 1u << (K - 1 - q);
    const uint32_t h = hm * hmin;
    const uint32_t step = n / (2 * h);
#pragma unroll
    for (uint32_t m = 0; m < R; m++) {
      if (m & hm) continue;
      Fr u = e[m], v = e[m + hm];
      e[m] = Fr::add(u, v);
      Fr d = Fr::sub(u, v);
      uint32_t ex = (off + (m & (hm - 1)) * hmin) * step;
      e[m + hm] = ex ? Fr::mul(d, tw[ex]) : d;
    }
  }
#pragma unroll
  for (uint32_t m = 0; m < R; m++) x[i0 + (size_t)m * hmin] = e[m];
}

// One DIT pass: K stages with half-spans h0, 2 h0, ..., h0 << (K-1).   (bit-reversed -> natural)
template <class Fr, int K>
ZKB_HDN inline void ntt_dit_body(Fr* x, const Fr* tw, uint32_t log_n, uint32_t h0, uint32_t t) {
  constexpr uint32_t R = 1u << K;
  const uint32_t n = 1u << log_n;
  if (t >= (n >> K)) return;
  const uint32_t off = t & (h0 - 1), blk = t / h0;
  const size_t i0 = (size_t)blk * ((size_t)R * h0) + off;
  Fr e[R];
#pragma unroll
  for (uint32_t m = 0; m < R; m++) e[m] = x[i0 + (size_t)m * h0];
#pragma unroll
  for (int q = 0; q < K; q++) {
    const uint32_t hm = 1u << q;
    const uint32_t h = hm * h0;
    const uint32_t step = n / (2 * h);
#pragma unroll
    for (uint32_t m = 0; m < R; m++) {
      if (m & hm) continue;
      uint32_t ex = (off + (m & (hm - 1)) * h0) * step;
      Fr u = e[m];
      Fr v = ex ? Fr::mul(e[m + hm], tw[ex]) : e[m + hm];
      e[m] = Fr::add(u, v);
      e[m + hm] = Fr::sub(u, v);
    }
  }
#pragma unroll
  for (uint32_t m = 0; m < R; m++) x[i0 + (size_t)m * h0] = e[m];
}

// x[i] *= table[bitrev(i)]  (coset shift / 1/n scaling applied to a bit-reversed coefficient vector)
template <class Fr>
ZKB_HDN inline void ntt_scale_brev_body(Fr* x, const Fr* table, uint32_t log_n, uint32_t t) {
  if (t >= (1u << log_n)) return;
  x[t] = Fr::mul(x[t], table[bitrev32(t, log_n)]);
}

// out[bitrev(i)] = in[i] * table[bitrev(i)] (table may be null), optionally leaving Montgomery form
template <class Fr>
ZKB_HDN inline void ntt_brev_copy_body(const Fr* in, Fr* out, const Fr* table, uint32_t log_n, int to_canonical, uint32_t t) {
  if (t >= (1u << log_n)) return;
  uint32_t j = bitrev32(t, log_n);
  Fr v = in[t];
  if (table) v = Fr::mul(v, table[j]);
  if (to_canonical) v = Fr::from_mont(v);
  out[j] = v;
}

// ab[i] = (a[i]*b[i] - c[i]) * zinv      (witness_map: a∘b - c, divided by the vanishing polynomial)
template <class Fr>
ZKB_HDN inline void qap_pointwise_body(Fr* a, const Fr* b, const Fr* c, Fr zinv, uint32_t n, uint32_t t) {
  if (t >= n) return;
  a[t] = Fr::mul(Fr::sub(Fr::mul(a[t], b[t]), c[t]), zinv);
}

// Montgomery conversion of a vector: dir = 0 canonical -> Montgomery, 1 Montgomery -> canonical
template <class Fr>
ZKB_HDN inline void fr_convert_body(const Fr* in, Fr* out, int dir, size_t n, size_t t) {
  if (t >= n) return;
  out[t] = dir ? Fr::from_mont(in[t]) : Fr::to_mont(in[t]);
}

// CSR sparse matrix-vector product in Montgomery form; rows >= n_rows are left to the caller.
template <class Fr>
ZKB_HDN inline void spmv_body(const uint32_t* rowptr, const uint32_t* col, const Fr* val, const Fr* z, Fr* out,
                              uint32_t n_rows, uint32_t t) {
  if (t >= n_rows) return;
  Fr acc = Fr::zero();
  for (uint32_t k = rowptr[t]; k < rowptr[t + 1]; k++) acc = Fr::add(acc, Fr::mul(val[k], z[col[k]]));
  out[t] = acc;
}

}  // namespace zkb
