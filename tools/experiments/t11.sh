timeout 900 python -m pytest tests/test_gpu_input_pipeline.py tests/test_gpu_variants.py -x -q > gpurun_out/r2n_pytest.txt 2>&1
