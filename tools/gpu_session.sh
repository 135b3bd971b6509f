# round-2 GPU session: parity tier, the bench line of both arms, launch list.  usage: gpurun -- bash tools/gpu_session.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_host.txt; lscpu | head -20 >> gpurun_out/${TAG}_host.txt; free -g >> gpurun_out/${TAG}_host.txt
(timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40) > gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -5 gpurun_out/${TAG}_pytest.log; head -c 400 gpurun_out/${TAG}_bench.json; echo; head -c 400 gpurun_out/${TAG}_bench_reference.json
