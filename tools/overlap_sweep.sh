#!/bin/bash
# overlap tuning sweep on N ranks (proxy regime: chain ~ cone): config 3 grid, reduced frame
# usage: overlap_sweep.sh <nproc> <width> <height> tune1 tune2 ...
N=$1; W=$2; H=$3; shift 3
port=29600
for t in "$@"; do
  port=$((port+1))
  tt=$t; [ "$t" == "base" ] && tt=""
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config 3 --mode shard --width $W --height $H --steps 40 --warmup 5 --tune "$tt" 2>/dev/null | python -c "
import sys, json
t=sys.stdin.read(); d=json.loads(t[t.index('{\"metric'):])
print('$t', 'fps', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'passes', {k:int(v) for k,v in d['passes_us_max_over_ranks'].items()})"
done
