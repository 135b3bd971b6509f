# 8-GPU: high-priority NCCL stream (default in bench.py now) and small CTA budget.  usage: gpurun --gpus 8 -- bash tools/gpu_session_multi4.sh 8 <tag>
N=$1; TAG=${2:-r2q}
mkdir -p gpurun_out
run() { port=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
run 29531 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_bench_g$N.json 2> gpurun_out/${TAG}_bench_g$N.err
ZKB_NCCL_MAX_CTAS=4 run 29532 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_bench_g${N}_ctas4.json 2>> gpurun_out/${TAG}_bench_g$N.err
ZKB_WM_SHARE=0 run 29533 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_bench_g${N}_replicated.json 2>> gpurun_out/${TAG}_bench_g$N.err
for f in gpurun_out/${TAG}_bench_g$N*.json; do python - "$f" <<'PY'
import json,sys
t=[l for l in open(sys.argv[1]) if l.startswith('{')]
if t:
    d=json.loads(t[-1]); print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],3), '%.3g'%d['value'], 'e2e', round(d['e2e']['ms_per_step'],3), d.get('latency_ms_one_proof_e2e'), json.dumps(d.get('stages_ms'))[:900])
else: print(sys.argv[1], 'NO JSON')
PY
done
tail -n 3 gpurun_out/${TAG}_bench_g$N.err | cut -c1-300
