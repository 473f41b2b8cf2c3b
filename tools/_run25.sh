mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/r02s_bench_n8.json 2> gpurun_out/r02s_bench_n8.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02s_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s_bench_n8.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
h=d['handoff_allgather']; print('strict', h['ms_per_step'], h['value'], h['handoff_checked'], h['transport']); print('split', json.dumps(h['split_phase'])[:400])
for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('value'), v.get('cropper'), v.get('crop_checked'))
PY
