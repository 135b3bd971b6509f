// BLS12-381 instantiation of the proving engine (384-bit base field, 12 x 32-bit limbs).
#include "engine.cuh"
#include "setup.cuh"
#include "gm17.cuh"
namespace zkb {
typedef Engine<CurveT<Bls381Fr, Bls381Fq>> EngineBls381;
EngineBase* make_engine_bls12_381(Stream st) { return new EngineBls381(st); }
size_t partial_bytes_bls12_381() { return sizeof(EngineBls381::Partial); }
}  // namespace zkb
