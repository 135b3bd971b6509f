// Host emulation harness (TEST ONLY): compiles the device headers as plain C++.
#include "fp.cuh"
using namespace zkb;
template <class P> static void do_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* o) {
  Fp<P> x, y, r;
  for (int i = 0; i < P::N; i++) { x.v[i] = a[i]; y.v[i] = b[i]; }
  switch (op) {
    case 0: r = Fp<P>::mul(x, y); break;
    case 1: r = Fp<P>::add(x, y); break;
    case 2: r = Fp<P>::sub(x, y); break;
    case 3: r = Fp<P>::neg(x); break;
    case 4: r = Fp<P>::inv(x); break;
    case 5: r = Fp<P>::to_mont(x); break;
    case 6: r = Fp<P>::from_mont(x); break;
    default: r = Fp<P>::zero();
  }
  for (int i = 0; i < P::N; i++) o[i] = r.v[i];
}
extern "C" void emu_fp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* o) {
  switch (field) {
    case 0: do_op<Bn254Fr>(op, a, b, o); break;
    case 1: do_op<Bn254Fq>(op, a, b, o); break;
    case 2: do_op<Bls381Fr>(op, a, b, o); break;
    case 3: do_op<Bls381Fq>(op, a, b, o); break;
  }
}
