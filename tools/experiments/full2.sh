timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.txt 2>&1
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
