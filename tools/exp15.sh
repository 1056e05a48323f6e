timeout 300 python tests/perf_conv.py 4096 2>&1 | head -6
