# round-2 session 6: full parity tier, plan-stream A/B, chunk-target sweep.   usage: gpurun -- bash tools/gpu_session6.sh <tag>
TAG=${1:-r2f}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --opt 12=0 --skip-cpu-baseline > gpurun_out/${TAG}_bench_planmain.json 2>> gpurun_out/${TAG}_bench.err
for t in 300000 150000; do
  timeout 200 python bench.py --steps 10 --warmup 3 --opt 13=$t --skip-cpu-baseline > gpurun_out/${TAG}_bench_chunk$t.json 2>> gpurun_out/${TAG}_bench.err
done
for c in 18 19; do
  timeout 200 python bench.py --steps 10 --warmup 3 --table-c $c --skip-cpu-baseline > gpurun_out/${TAG}_bench_c$c.json 2>> gpurun_out/${TAG}_bench.err
done
timeout 200 python bench.py --steps 10 --warmup 3 --witness bits --skip-cpu-baseline > gpurun_out/${TAG}_bench_bits.json 2>> gpurun_out/${TAG}_bench.err
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('latency_ms_one_proof_e2e'), (d.get('tables') or {}).get('c_z'), json.dumps(d.get('stages_ms')))
"; done
tail -n 5 gpurun_out/${TAG}_bench.err
