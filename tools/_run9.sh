mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize.py > gpurun_out/r02_sanitizer_synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_synccheck.txt
export PCL_BENCH_NO_GRAPH=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"render_kernel_ring|crop_kernel" -s 2 -c 4 -f -o gpurun_out/r02i_render_crop python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_render.log 2>&1
ls -la gpurun_out | tail -8
