#!/bin/bash
# Developer tool (GPU box with >= 2 GPUs): what the data-parallel step costs over the single-GPU one at 16 patches per GPU --
# SyncBatchNorm exchange on / off (per-rank statistics) x gradient all-reduce on / off (SRG_DP_NO_ALLREDUCE=1: measurement only).
# usage: tools/apportion_dp.sh [N]    prints one line per combination
N=${1:-2}
cd "$(dirname "$0")/.."
for cfg in "peer 0" "none 0" "peer 1" "none 1"; do
  set -- $cfg
  SRG_DP_NO_ALLREDUCE=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) \
    bench.py --gpus $N --batch 16 --steps 30 --warmup 5 --no-hbm --no-e2e --syncbn $1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('gpus $N syncbn=$1 no_allreduce=$2:', round(d['ms_per_step'], 3), 'ms per step,', round(d['value'], 1), 'patches/s, clk', d['clocks']['sm_mhz'])
"
done
python bench.py --steps 30 --warmup 5 --no-hbm --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('gpus 1:', round(d['ms_per_step'], 3), 'ms per step,', round(d['value'], 1), 'patches/s')
"
