# round-2 session 5: parity tier, batch-affine A/B (rounds 0 / 2 / 3 / 4), window-width sweep with the affine rounds, launch list.
# usage: gpurun -- bash tools/gpu_session5.sh <tag>
TAG=${1:-r2e}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
for r in 0 2 4; do
  timeout 200 python bench.py --steps 10 --warmup 3 --opt 10=$r --skip-cpu-baseline > gpurun_out/${TAG}_bench_ba$r.json 2>> gpurun_out/${TAG}_bench.err
done
for c in 15 16 18 19; do
  timeout 200 python bench.py --steps 10 --warmup 3 --table-c $c --skip-cpu-baseline > gpurun_out/${TAG}_bench_c$c.json 2>> gpurun_out/${TAG}_bench.err
done
timeout 200 python bench.py --steps 10 --warmup 3 --witness bits --skip-cpu-baseline > gpurun_out/${TAG}_bench_bits.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prove_loop.py 20 2 > gpurun_out/${TAG}_launches.log 2>&1
ZKB_CPU_MAX=0 timeout 300 python tools/microbench.py 20 22 > gpurun_out/${TAG}_microbench.log 2>&1
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('latency_ms_one_proof_e2e'), (d.get('tables') or {}).get('c_z'), json.dumps(d.get('stages_ms')))
"; done
tail -3 gpurun_out/${TAG}_microbench.log | cut -c1-300; tail -n 5 gpurun_out/${TAG}_bench.err
