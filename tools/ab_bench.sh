# A/B harness: same box, same process conditions; prints ms/step for each environment setting (developer tool)
# usage: tools/ab_bench.sh "VAR=1 OTHER=2" "VAR=0" ...   (REPS=n repetitions, default 2)
for rep in $(seq 1 ${REPS:-2}); do
for cfg in "$@"; do
  env $cfg python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-hbm 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'avg_launch_us', round(d['roofline']['avg_launch_us'],2), 'clk', d['clocks']['sm_mhz'])
"
done; done
