mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "render or Render or layers" 2>&1 | tail -5
timeout 900 python tools/render_ab.py 1 4 2>&1 | tee gpurun_out/r02c_render_ab.txt
timeout 300 python bench.py --no-configs > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1c.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['render_roofline'])
PY
export PCL_BENCH_NO_GRAPH=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render -s 3 -c 1 -f -o gpurun_out/r02c_render python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/ncu_render.log 2>&1
