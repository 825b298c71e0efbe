export PPM_FORCE_BAND=1
timeout 300 python tools/run_band.py --genomes 10000 --caterpillar --genes 24 --params 1 --seed 41000 --variants default,noband,V=1 --reps 1 > gpurun_out/r2o_cat24.txt 2>&1
timeout 300 python tools/run_band.py --genomes 10000 --caterpillar --genes 3000 --params 2 --variants default,noband --reps 1 > gpurun_out/r2o_cat3000.txt 2>&1
timeout 300 python tools/run_band.py --genomes 8000 --caterpillar --genes 24 --params 1 --variants default,noband --reps 1 > gpurun_out/r2o_cat8000.txt 2>&1
