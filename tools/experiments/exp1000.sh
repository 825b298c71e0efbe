# usage: bash tools/experiments/exp1000.sh "ENV=.." ...   -- sweep time / algorithmic TFLOP/s of the 1,000-genome configuration
if [ $# -eq 0 ]; then set -- "" "PPM_NO_DEEP_STAGE=1" "PPM_R=8"; fi
for v in "$@"; do
  echo "== $v"; env $v python tools/run_config.py --genomes ${GENOMES:-1000} --genes ${GENES:-20000} --params ${PARAMS:-8} --reps 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['sweep_s'], d['tflops_algorithmic'], d['ll0'])"
done
