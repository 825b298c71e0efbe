set -x
export PPM_PLAN_DEBUG=1
timeout 600 python -m pytest tests/test_gpu_band.py -x -q > gpurun_out/r2c_pytest.txt 2>&1
timeout 300 python tools/run_band.py --genomes 5000 --genes 4000 --params 4 --variants default,V=1,G=4,G=16,G=40 --reps 2 > gpurun_out/r2c_n5000.txt 2>&1
PPM_FORCE_BAND=1 timeout 300 python tools/run_band.py --genomes 1000 --genes 20000 --params 32 --variants default,G=1,G=4,noband --reps 2 > gpurun_out/r2c_n1000.txt 2>&1
PPM_FORCE_BAND=1 timeout 300 python tools/run_band.py --genomes 10000 --caterpillar --genes 3000 --params 2 --variants default,noband --reps 2 > gpurun_out/r2c_cat.txt 2>&1
