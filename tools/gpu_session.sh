mkdir -p gpurun_out
(timeout 520 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/s4_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
ZKB_WM_EARLY=1 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/s4_bench_wmearly.json 2>> gpurun_out/s4_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --witness bits > gpurun_out/s4_bench_bits.json 2>> gpurun_out/s4_bench.err
timeout 200 python tools/shard_scan.py 20 > gpurun_out/s4_shard.jsonl 2>&1
for lg in 16 18 22; do timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --log-n $lg > gpurun_out/s4_bench_lg$lg.json 2>> gpurun_out/s4_bench.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4_launches.csv python tools/prove_loop.py 20 2 > gpurun_out/s4_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msm_accum1 -c 5 -f -o gpurun_out/s4_accum1 python tools/prove_loop.py 20 1 > gpurun_out/s4_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_d.._tile -c 4 -f -o gpurun_out/s4_ntt python tools/prove_loop.py 20 1 > gpurun_out/s4_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msm_bitsum -c 5 -f -o gpurun_out/s4_bitsum python tools/prove_loop.py 20 1 > gpurun_out/s4_ncu_c.log 2>&1
ZKB_CURVE=bls12_381 timeout 300 python tools/gpu_profile.py 20 > gpurun_out/s4_bls.log 2>&1
cat gpurun_out/s4_pytest.log; ls -la gpurun_out
