// BN254 (bn128) instantiation of the proving engine — separate translation unit so the two curves compile in parallel.
#include "engine.cuh"
#include "setup.cuh"
#include "gm17.cuh"
namespace zkb {
typedef Engine<CurveT<Bn254Fr, Bn254Fq>> EngineBn254;
EngineBase* make_engine_bn254(Stream st) { return new EngineBn254(st); }
size_t partial_bytes_bn254() { return sizeof(EngineBn254::Partial); }
}  // namespace zkb
