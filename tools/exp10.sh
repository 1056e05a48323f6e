python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t10.log; tail -3 gpurun_out/r2_t10.log
python tools/train_bench.py 512 10 bf16 attn 2>&1 | tail -1
SPDM_GN_BWD_T=256 python tools/train_bench.py 512 10 bf16 attn 2>&1 | tail -1
SPDM_GN_BWD_T=256 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
