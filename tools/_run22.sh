mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider tests/test_dist.py 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-configs > gpurun_out/r02r_bench_n2.json 2> gpurun_out/r02r_bench_n2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02r_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_bench_n2.json').read().strip().splitlines()[-1])
h=d['handoff_allgather']; print('strict', h['ms_per_step'], h['value'], h['handoff_checked']); print('split', json.dumps(h['split_phase'])[:500])
PY
