timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider tests/test_dist.py 2>&1 | tail -3
for x in 1 0; do
PCL_BENCH_HANDOFF_SINGLE_KERNEL=$x timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-configs > gpurun_out/hs$x.json 2> gpurun_out/hs$x.err; echo "bench rc=$?"; tail -c 200 gpurun_out/hs$x.err
python - <<PY
import json
d=json.loads(open('gpurun_out/hs$x.json').read().strip().splitlines()[-1])
h=d['handoff_allgather']; print('single_kernel=$x strict', round(h['ms_per_step']*1e3,2), h['handoff_checked'], 'split', json.dumps({k:h['split_phase'].get(k) for k in ('ms_per_step','handoff_checked','error')}))
PY
done
