mkdir -p gpurun_out
(timeout 520 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s6_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
ZKB_PRECOMP_C=19 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/s6_bench_c19.json 2>> gpurun_out/s6_bench.err
ZKB_PRECOMP_C=18 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/s6_bench_c18.json 2>> gpurun_out/s6_bench.err
timeout 200 python tools/shard_scan.py 20 > gpurun_out/s6_shard.jsonl 2>&1
mkdir -p /tmp/ncu
for spec in accum1:k_msm_accum1:5 ntt:k_ntt_d.._tile:4 bitsum:k_msm_bitsum:6; do
  name=${spec%%:*}; rest=${spec#*:}; pat=${rest%%:*}; cnt=${rest#*:}
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$pat -c $cnt -f -o /tmp/ncu/s6_$name python tools/prove_loop.py 20 1 > gpurun_out/s6_ncu_$name.log 2>&1
  ncu -i /tmp/ncu/s6_$name.ncu-rep --page raw --csv > gpurun_out/s6_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/s6_$name.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/s6_${name}_source.csv
done
cat gpurun_out/s6_pytest.log; ls -la gpurun_out | tail -18; du -sh gpurun_out
