mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -4 gpurun_out/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 50 python tools/sanitize.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --print-limit 400 python tools/sanitize.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck --print-limit 100 python tools/sanitize.py > gpurun_out/r02_sanitizer_synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_synccheck.txt
timeout 400 python bench.py > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1d.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['render_roofline']['frac'], d['render_roofline'].get('one_launch_over_all_batches'))
for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('value'))
PY
export PCL_BENCH_NO_GRAPH=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel_ring -s 2 -c 1 -f -o gpurun_out/r02j_render_ring python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/ncu_render.log 2>&1
