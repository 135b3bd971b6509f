# round-2 tuning session: parity tier, bench variants (bit-sum radix, window width), launch list + ncu of the accumulate kernels.
# usage: gpurun -- bash tools/gpu_session3.sh <tag>
TAG=${1:-r2c}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --opt 7=8 --skip-cpu-baseline > gpurun_out/${TAG}_bench_radix8.json 2>> gpurun_out/${TAG}_bench.err
for c in 17 18 19 20; do
  timeout 200 python bench.py --steps 10 --warmup 3 --table-c $c --skip-cpu-baseline > gpurun_out/${TAG}_bench_c$c.json 2>> gpurun_out/${TAG}_bench.err
done
timeout 200 python bench.py --steps 10 --warmup 3 --witness bits --skip-cpu-baseline > gpurun_out/${TAG}_bench_bits.json 2>> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --pipeline 1 --skip-cpu-baseline > gpurun_out/${TAG}_bench_p1.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prove_loop.py 20 2 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_msm_accum1 -c 10 -o gpurun_out/${TAG}_accum1 python tools/prove_loop.py 20 2 > gpurun_out/${TAG}_ncu_accum1.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
for f in gpurun_out/${TAG}_bench*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('latency_ms_one_proof_e2e'), (d.get('tables') or {}).get('c_z'), json.dumps(d.get('stages_ms')))
"; done
tail -5 gpurun_out/${TAG}_bench.err
