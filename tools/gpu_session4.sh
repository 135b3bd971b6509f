# round-2 session 4: parity tier (incl. the native front door), NTT kernel A/B, config 5 sweep on one GPU, ingestion / CLI wall time,
# config 4 on one GPU, ncu of the new NTT tile pass.   usage: gpurun -- bash tools/gpu_session4.sh <tag>
TAG=${1:-r2d}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --opt 9=1 --skip-cpu-baseline > gpurun_out/${TAG}_bench_ntt1.json 2>> gpurun_out/${TAG}_bench.err
ZKB_CPU_MAX=22 timeout 900 python tools/microbench.py 18 20 22 24 25 26 > gpurun_out/${TAG}_microbench.log 2>&1; cp gpurun_out/microbench.json gpurun_out/${TAG}_microbench.json 2>/dev/null
timeout 600 python tools/ingest_bench.py --log-n 20 > gpurun_out/${TAG}_ingest.json 2> gpurun_out/${TAG}_ingest.err
timeout 600 python bench.py --curve bls12_381 --log-n 22 --steps 4 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_bench_bls_2p22.json 2> gpurun_out/${TAG}_bench_bls.err
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:zkb_ntt_tile2 -c 6 -o gpurun_out/${TAG}_ntt python tools/prove_loop.py 20 1 > gpurun_out/${TAG}_ncu_ntt.log 2>&1
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('latency_ms_one_proof_e2e'), (d.get('tables') or {}), json.dumps(d.get('stages_ms')))
"; done
tail -8 gpurun_out/${TAG}_microbench.log | cut -c1-400; cat gpurun_out/${TAG}_ingest.json | cut -c1-1800; tail -3 gpurun_out/${TAG}_ingest.err gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_bench_bls.err
