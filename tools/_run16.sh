timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "crop or Crop or kats or better or story" 2>&1 | tail -2
timeout 300 python tools/handoff_probe.py 2>&1 | tee gpurun_out/r02p_handoff_probe.txt
