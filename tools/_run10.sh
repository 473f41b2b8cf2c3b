mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log); tail -6 gpurun_out/pytest_gpu.log
for v in 7 8; do
PCL_SCROLLY_VARIANT=$v timeout 300 python bench.py --no-configs > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench v$v rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v$v.json').read().strip().splitlines()[-1])
print('v$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_checked']['envs'], d['env_errors'])
PY
done
