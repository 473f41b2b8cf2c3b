mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -8 gpurun_out/pytest_gpu.log
(timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1); cat gpurun_out/e2e_probe.txt | tail -15
export PCL_BENCH_NO_GRAPH=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scrolly_maze_step -s 30 -c 1 -f -o gpurun_out/r02a_step python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render -s 3 -c 2 -f -o gpurun_out/r02a_render python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_render.log 2>&1
ls -la gpurun_out
