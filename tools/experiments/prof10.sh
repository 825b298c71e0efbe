timeout 900 python -m pytest tests/test_gpu_band.py tests/test_gpu_parity.py -x -q > gpurun_out/r2l_pytest.txt 2>&1
timeout 300 python tools/run_band.py --genomes 5000 --genes 20000 --params 4 --variants default,SCRATCH_MB=2048,SCRATCH_MB=8192 --reps 2 > gpurun_out/r2l_n5000.txt 2>&1
timeout 300 python tools/run_band.py --genomes 5000 --genes 20000 --params 1 --variants default --reps 2 > gpurun_out/r2l_n5000p1.txt 2>&1
