mkdir -p gpurun_out
./tools/probes/pdl_probe 2>&1 | tee gpurun_out/r02f_pdl_probe.txt
timeout 900 python tools/render_ab.py sweep 1 5 7 2>&1 | tee gpurun_out/r02f_render_sweep.txt
