timeout 900 python -m pytest tests/test_gpu_band.py tests/test_gpu_parity.py -x -q > gpurun_out/r2j_pytest.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_launches5000.csv python tools/run_band.py --genomes 5000 --genes 4000 --params 1 --variants default --reps 1 > /dev/null 2>&1
timeout 300 python tools/run_band.py --genomes 5000 --genes 20000 --params 4 --variants default --reps 2 > gpurun_out/r2j_n5000.txt 2>&1
