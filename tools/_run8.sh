mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02h_bench_n8.json 2> gpurun_out/r02h_bench_n8.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02h_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench_n8.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','per_rank_ms_per_step','e2e','handoff_allgather','numa','clocks'): print(k, json.dumps(d.get(k))[:1200])
for k,v in (d.get('configs') or {}).items(): print(k, json.dumps({a:b for a,b in v.items() if a in ('value','ms_per_step','e2e','global_batch')})[:500])
PY
