timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.txt 2>&1
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
