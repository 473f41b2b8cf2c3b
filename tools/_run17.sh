for v in A B A B; do echo "== lib $v"; PCL_LIB_PATH=$PWD/_ab/libpcl_$v.so timeout 300 python tools/handoff_probe.py 2>&1 | tail -3; done | tee gpurun_out/r02q_handoff_ab.txt
