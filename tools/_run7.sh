mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider tests/test_dist.py 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02g_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','per_rank_ms_per_step','e2e','handoff_allgather','numa','parity_checked'): print(k, json.dumps(d.get(k))[:900])
PY
