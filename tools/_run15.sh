mkdir -p gpurun_out
timeout 300 python tools/handoff_probe.py 2>&1 | tee gpurun_out/r02o_handoff_probe.txt
PCL_BENCH_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crop_handoff_kernel -s 20 -c 1 -f -o gpurun_out/r02o_crop_handoff python tools/handoff_probe.py > gpurun_out/ncu_handoff.log 2>&1; tail -3 gpurun_out/ncu_handoff.log
