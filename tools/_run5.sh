mkdir -p gpurun_out
for v in 4 5 7; do PCL_RENDER_VARIANT=$v timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "render or Render or layers" 2>&1 | tail -2; done
timeout 900 python tools/render_ab.py 1 4 5 7 7:1 7:3 2>&1 | tee gpurun_out/r02e_render_ab.txt
