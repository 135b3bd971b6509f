mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s10_pytest.log
timeout 120 python bench.py --steps 10 --warmup 3 > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
timeout 100 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --witness bits > gpurun_out/s10_bench_bits.json 2>> gpurun_out/s10_bench.err
timeout 100 python tools/shard_scan.py 20 4 8 > gpurun_out/s10_shard.jsonl 2>&1
cat gpurun_out/s10_pytest.log; head -c 300 gpurun_out/s10_bench.json
