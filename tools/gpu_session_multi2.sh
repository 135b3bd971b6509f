# round-2 multi-GPU session.  usage: gpurun --gpus N -- bash tools/gpu_session_multi2.sh N <tag>
N=$1; TAG=${2:-r2m}
mkdir -p gpurun_out
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
run 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_g$N.json 2> gpurun_out/${TAG}_bench_g$N.err
run 29512 bench.py --gpus $N --steps 10 --warmup 3 --opt 7=8 --skip-cpu-baseline > gpurun_out/${TAG}_bench_g${N}_radix8.json 2>> gpurun_out/${TAG}_bench_g$N.err
if [ $N -ge 8 ]; then
  run 29513 bench.py --gpus $N --steps 5 --warmup 3 --curve bls12_381 --log-n 22 --skip-cpu-baseline > gpurun_out/${TAG}_bench_bls_2p22_g$N.json 2> gpurun_out/${TAG}_bench_bls_g$N.err
  run 29514 tools/microbench_multi.py 20 22 24 26 > gpurun_out/${TAG}_microbench_g$N.jsonl 2> gpurun_out/${TAG}_microbench_g$N.err
else
  run 29514 tools/microbench_multi.py 20 22 24 > gpurun_out/${TAG}_microbench_g$N.jsonl 2> gpurun_out/${TAG}_microbench_g$N.err
fi
for f in gpurun_out/${TAG}_bench*g$N*.json; do python - "$f" <<'PY'
import json,sys
t=[l for l in open(sys.argv[1]) if l.startswith('{')]
if t:
    d=json.loads(t[-1]); print(sys.argv[1], d['n_gpus'], round(d['ms_per_step'],3), '%.3g'%d['value'], 'e2e', round(d['e2e']['ms_per_step'],3), d.get('latency_ms_one_proof_e2e'), json.dumps(d.get('stages_ms'))[:700])
else: print(sys.argv[1], 'NO JSON')
PY
done
cat gpurun_out/${TAG}_microbench_g$N.jsonl | cut -c1-300; tail -n 4 gpurun_out/${TAG}_bench_g$N.err gpurun_out/${TAG}_microbench_g$N.err | cut -c1-300
