# pipelining + window-width sweep.  usage: gpurun -- bash tools/gpu_session2.sh <tag>
TAG=${1:-r2b}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --pipeline 1 --skip-cpu-baseline > gpurun_out/${TAG}_bench_p1.json 2>> gpurun_out/${TAG}_bench.err
for c in 18 19 20 21; do
  timeout 200 python bench.py --steps 10 --warmup 3 --table-c $c --skip-cpu-baseline > gpurun_out/${TAG}_bench_c$c.json 2>> gpurun_out/${TAG}_bench.err
done
timeout 200 python bench.py --steps 10 --warmup 3 --witness bits --skip-cpu-baseline > gpurun_out/${TAG}_bench_bits.json 2>> gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -3 gpurun_out/${TAG}_pytest.log
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('latency_ms_one_proof_e2e'), (d.get('tables') or {}).get('c_z'), (d.get('cpu_baseline') or {}).get('cores'))
"; done
tail -5 gpurun_out/${TAG}_bench.err
