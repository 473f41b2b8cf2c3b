mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider tests/test_dist.py -k "two_gpus or handoff or crop" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-configs > gpurun_out/r02n_bench_n2.json 2> gpurun_out/r02n_bench_n2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02n_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02n_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','handoff_allgather'): print(k, json.dumps(d.get(k))[:700])
PY
