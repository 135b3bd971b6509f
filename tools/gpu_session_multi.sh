# usage: bash tools/gpu_session_multi.sh N   — N-GPU bench (NCCL), with and without the shared witness map
N=$1
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$2 bench.py --gpus $N --steps 10 --warmup 3 "${@:3}"; }
run 1 > gpurun_out/s7_bench_g$N.json 2> gpurun_out/s7_bench_g$N.err
if [ $N -ge 3 ]; then ZKB_WM_SHARE=0 run 2 > gpurun_out/s7_bench_g${N}_replicated.json 2>> gpurun_out/s7_bench_g$N.err; fi
tail -3 gpurun_out/s7_bench_g$N.err | cut -c1-300
for f in gpurun_out/s7_bench_g$N*.json; do python - "$f" <<'PY'
import json,sys
t=[l for l in open(sys.argv[1]) if l.startswith('{')]
if t:
    d=json.loads(t[-1]); print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],3), '%.3g'%d['value'], 'e2e', round(d['e2e']['ms_per_step'],3), d['config']['parallelism'])
else: print(sys.argv[1], 'NO JSON')
PY
done
