python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
python tools/train_bench.py 512 10 bf16 attn 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:gn_bwd_cached --launch-skip 64 -c 6 -o gpurun_out/r02_gn_bwd -f python tools/train_bench.py 512 1 bf16 attn > gpurun_out/ncu_gn.log 2>&1; tail -2 gpurun_out/ncu_gn.log
