timeout 600 python -m pytest tests/test_apprehend.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "apprehend or marauders" 2>&1 | tail -15
