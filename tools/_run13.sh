timeout 900 python tools/render_ab.py sweep 1 5 2>&1 | tee gpurun_out/r02m_render_sweep.txt
