# round-2 final single-GPU session: whole parity tier, both bench arms, G2 occupancy variant, ingestion / trait-shaped call.
# usage: gpurun -- bash tools/gpu_session_final.sh <tag>
TAG=${1:-r2z}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -30) > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
if [ -f zokrates_b200/libzkb200_g2m3.so ]; then
  ZKB200_LIB=$PWD/zokrates_b200/libzkb200_g2m3.so timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_bench_g2minb3.json 2>> gpurun_out/${TAG}_bench.err
fi
timeout 200 python bench.py --steps 10 --warmup 3 --witness bits --skip-cpu-baseline > gpurun_out/${TAG}_bench_bits.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python tools/ingest_bench.py --log-n 20 > gpurun_out/${TAG}_ingest.json 2> gpurun_out/${TAG}_ingest.err
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('latency_ms_one_proof_e2e'), json.dumps(d.get('trait_shaped_call')), json.dumps(d.get('roofline'))[:600], json.dumps(d.get('stages_ms'))[:500])
"; done
cut -c1-1500 gpurun_out/${TAG}_ingest.json; tail -n 4 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_ingest.err
