echo PAIR=0; SPDM_PAIR=0 timeout 120 python tests/perf_conv.py 4096 2>&1 | sed -n 4,6p
echo PAIR=1; SPDM_PAIR=1 timeout 120 python tests/perf_conv.py 4096 2>&1 | sed -n 4,5p
