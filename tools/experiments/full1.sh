timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
