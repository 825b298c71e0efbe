timeout 900 python -m pytest tests/test_gpu_band.py -x -q -k "sparse or stream or golden" > gpurun_out/r2p_pytest.txt 2>&1
timeout 300 python tools/run_band.py --genomes 5000 --genes 20000 --params 4 --variants default,sparse --reps 2 > gpurun_out/r2p_n5000.txt 2>&1
PPM_FORCE_BAND=1 timeout 300 python tools/run_band.py --genomes 1000 --genes 20000 --params 32 --variants default,sparse --reps 2 > gpurun_out/r2p_n1000.txt 2>&1
