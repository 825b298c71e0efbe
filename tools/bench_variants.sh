# usage: bash tools/bench_variants.sh "label:ENV=1 ENV2=2" ...   (default: a standard set)
run() { env $2 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['fp64_pipe_frac'])"; }
if [ $# -eq 0 ]; then set -- "default:" "ni2:PPM_NI=2" "tmem12:PPM_FORCE_TMEM=1 PPM_TMEM12=1"; fi
for v in "$@"; do run "${v%%:*}" "${v#*:}"; done
