export PPM_PLAN_DEBUG=1
ncu --set full --import-source on --clock-control none -k regex:band_integral -c 1 -f -o gpurun_out/r2g_int5000 python tools/run_band.py --genomes 5000 --genes 4000 --params 4 --variants default --reps 1 > gpurun_out/r2g_ncu5000.log 2>&1
PPM_FORCE_BAND=1 ncu --set full --import-source on --clock-control none -k regex:band_integral -c 1 -f -o gpurun_out/r2g_int1000 python tools/run_band.py --genomes 1000 --genes 20000 --params 32 --variants default --reps 1 > gpurun_out/r2g_ncu1000.log 2>&1
