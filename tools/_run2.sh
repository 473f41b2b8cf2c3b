mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "render or Render or layers" 2>&1 | tail -5
timeout 900 python tools/render_ab.py 1 2 3 4 4:2 4:3 4:8 2>&1 | tee gpurun_out/r02b_render_ab.txt
timeout 300 python bench.py --no-configs > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1b.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['render_roofline'])
PY
