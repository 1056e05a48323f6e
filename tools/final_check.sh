python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_final_tests.log; tail -3 gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -4 gpurun_out/r2_final_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 300 gpurun_out/r2_final_bench.err
