run() { python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['fp64_pipe_frac'])"; }
run default
PPM_FORCE_TMEM=1 PPM_TMEM12=1 run tmem12
PPM_FORCE_TMEM=1 run tmem16
PPM_R=12 run r12
PPM_R=16 run r16
