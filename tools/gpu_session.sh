mkdir -p gpurun_out
(timeout 520 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s6_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
ZKB_PRECOMP_C=19 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/s6_bench_c19.json 2>> gpurun_out/s6_bench.err
ZKB_PRECOMP_C=18 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/s6_bench_c18.json 2>> gpurun_out/s6_bench.err
timeout 200 python tools/shard_scan.py 20 > gpurun_out/s6_shard.jsonl 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_msm_accum1 -c 5 -f -o gpurun_out/s6_accum1 python tools/prove_loop.py 20 1 > gpurun_out/s6_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_ntt_d.._tile -c 4 -f -o gpurun_out/s6_ntt python tools/prove_loop.py 20 1 > gpurun_out/s6_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_msm_bitsum -c 5 -f -o gpurun_out/s6_bitsum python tools/prove_loop.py 20 1 > gpurun_out/s6_ncu_c.log 2>&1
cat gpurun_out/s6_pytest.log; ls -la gpurun_out | tail -15
