mkdir -p gpurun_out
(timeout 520 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s9_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
cat gpurun_out/s9_pytest.log; head -c 300 gpurun_out/s9_bench.json
