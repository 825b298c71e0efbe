set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_plain_configs3.json 2> gpurun_out/r02_plain_configs3.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_configs3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:band_integral -s 2 -c 1 -f -o gpurun_out/r02_integral_configs3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_ncu_full3.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --workload "configs[4]" > gpurun_out/r02_plain_configs4.json 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:band_integral -s 1 -c 1 -f -o gpurun_out/r02_integral_configs4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --workload "configs[4]" > gpurun_out/r02_ncu_full4.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --workload "configs[1]" > gpurun_out/r02_plain_configs1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_configs1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --workload "configs[1]" > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:sweep_kernel -s 3 -c 1 -f -o gpurun_out/r02_sweep_configs1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --workload "configs[1]" > gpurun_out/r02_ncu_full1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:sweep_kernel -s 1 -c 1 -f -o gpurun_out/r02_sweep_configs2_P32 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --workload "configs[2]" --params 32 > gpurun_out/r02_ncu_full2.log 2>&1
