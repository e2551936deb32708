N=${1:-2}
STEPS=${2:-60}
[ "$N" = "2" ] && python -m pytest tests -x -q -m gpu -k "fused_sharded or sparse_mip" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps $STEPS --warmup 10"
run() { echo "== $*"; $TR "$@" > /tmp/bench_out.txt 2>&1; grep -m3 "Error" /tmp/bench_out.txt; tail -1 /tmp/bench_out.txt | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
  j=json.loads(l); print(j['value'], j['unit'], j['ms_per_step'], 'e2e', j.get('e2e',{}).get('value')); print('  rank0', j.get('passes_us')); print('  max  ', j.get('passes_us_max_over_ranks'))
except Exception as e: print('ERR', l[-1500:])
"; }
run --config 3 --mode shard
run --config 3 --mode tiles
run --config 4 --mode shard
run --config 2 --mode shard
[ "$N" != "2" ] && run --config 2 --mode views
