mkdir -p gpurun_out
PCL_RENDER_VARIANT=5 timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "render or Render or layers" 2>&1 | tail -2
timeout 900 python tools/render_ab.py 1 5 5:3 5:2 5:1 2>&1 | tee gpurun_out/r02l_render_ab.txt
export PCL_BENCH_NO_GRAPH=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 2 -c 1 -f -o gpurun_out/r02l_render python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/ncu_render.log 2>&1
grep -c render_kernel_ring gpurun_out/ncu_render.log
