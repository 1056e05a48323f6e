python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || tail -c 400 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err || tail -c 300 gpurun_out/r02_bench_reference.err
python tools/sweep.py > gpurun_out/r02_sweep_1gpu.md 2> gpurun_out/r02_sweep_1gpu.err || tail -c 300 gpurun_out/r02_sweep_1gpu.err
tail -3 gpurun_out/r02_sweep_1gpu.md | cut -c1-200
python tools/margins.py 2>&1 | tail -6
