set -x
export PPM_PLAN_DEBUG=1
ncu --set full --import-source on --clock-control none -k band_kernel -c 1 -f -o gpurun_out/r2b_band5000 python tools/run_band.py --genomes 5000 --genes 4000 --params 4 --variants default --reps 1 > gpurun_out/r2b_ncu5000.log 2>&1
PPM_FORCE_BAND=1 python tools/run_band.py --genomes 1000 --genes 20000 --params 32 --variants default --reps 2 > gpurun_out/r2b_n1000.txt 2>&1
PPM_FORCE_BAND=1 ncu --set full --import-source on --clock-control none -k band_kernel -c 1 -f -o gpurun_out/r2b_band1000 python tools/run_band.py --genomes 1000 --genes 20000 --params 32 --variants default --reps 1 > gpurun_out/r2b_ncu1000.log 2>&1
