set -x
export PPM_PLAN_DEBUG=1
timeout 600 python -m pytest tests/test_gpu_band.py -x -q > gpurun_out/r2e_pytest.txt 2>&1
timeout 300 python tools/run_band.py --genomes 5000 --genes 4000 --params 4 --variants default,G=16 --reps 2 > gpurun_out/r2e_n5000.txt 2>&1
PPM_FORCE_BAND=1 timeout 300 python tools/run_band.py --genomes 1000 --genes 20000 --params 32 --variants default,noband --reps 2 > gpurun_out/r2e_n1000.txt 2>&1
PPM_FORCE_BAND=1 timeout 300 python tools/run_band.py --genomes 10000 --caterpillar --genes 3000 --params 2 --variants default --reps 2 > gpurun_out/r2e_cat.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:band_integral -c 1 -f -o gpurun_out/r2e_int5000 python tools/run_band.py --genomes 5000 --genes 4000 --params 4 --variants default --reps 1 > gpurun_out/r2e_ncu5000.log 2>&1
