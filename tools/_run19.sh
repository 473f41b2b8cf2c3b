timeout 600 python -m pytest tests/test_shockwave.py tests/test_apprehend.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25
