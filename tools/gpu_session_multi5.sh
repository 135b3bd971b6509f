# 8-GPU: chain-owner share A/B.  usage: gpurun --gpus 8 -- bash tools/gpu_session_multi5.sh 8 <tag>
N=$1; TAG=${2:-r2u}
mkdir -p gpurun_out
run() { port=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
run 29541 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_bench_g$N.json 2> gpurun_out/${TAG}_bench_g$N.err
run 29542 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline --opt 14=70 > gpurun_out/${TAG}_bench_g${N}_share70.json 2>> gpurun_out/${TAG}_bench_g$N.err
for f in gpurun_out/${TAG}_bench_g$N*.json; do python - "$f" <<'PY'
import json,sys
t=[l for l in open(sys.argv[1]) if l.startswith('{')]
if t:
    d=json.loads(t[-1]); print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],3), '%.3g'%d['value'], 'e2e', round(d['e2e']['ms_per_step'],3), d.get('latency_ms_one_proof_e2e'), json.dumps(d.get('stages_ms'))[:900])
else: print(sys.argv[1], 'NO JSON')
PY
done
tail -n 3 gpurun_out/${TAG}_bench_g$N.err | cut -c1-300
