mkdir -p gpurun_out
(timeout 520 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s8_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s8_bench_reference.json 2>> gpurun_out/s8_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --witness bits > gpurun_out/s8_bench_bits.json 2>> gpurun_out/s8_bench.err
timeout 200 python tools/shard_scan.py 20 > gpurun_out/s8_shard.jsonl 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_smoke.log 2>&1
cat gpurun_out/s8_pytest.log; tail -2 gpurun_out/s8_smoke.log; head -c 400 gpurun_out/s8_bench.json
