mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
for m in 0 1; do
PCL_BENCH_SEPARATE_CROP=$m timeout 400 python bench.py > gpurun_out/bench_crop$m.json 2> gpurun_out/bench_crop$m.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_crop$m.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_crop$m.json').read().strip().splitlines()[-1])
print('separate=$m', d['value'], d['ms_per_step'], d['e2e']['value'])
v=d['configs']['C5_scrolly64_crop9']; print(v['value'], v['ms_per_step'], v['e2e']['value'], v.get('cropper'), v.get('crop_checked'), v['launches_per_step'])
PY
done
