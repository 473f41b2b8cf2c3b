for x in 0 1 2; do
PCL_HANDOFF_EXPERIMENT=$x timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-configs > gpurun_out/hx$x.json 2> gpurun_out/hx$x.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/hx$x.json').read().strip().splitlines()[-1])
h=d['handoff_allgather']; print('exp $x strict', round(h['ms_per_step']*1e3,2), 'split', json.dumps({k:h['split_phase'].get(k) for k in ('ms_per_step','handoff_checked','error')}))
PY
done
