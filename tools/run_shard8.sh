N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 100 --warmup 10"
run() { echo "== $*"; $TR "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
  j=json.loads(l); print(j['value'], j['unit'], j['ms_per_step'], j.get('e2e',{}).get('value'), j.get('passes_us'))
except Exception as e: print('ERR', l[-1500:])
"; }
run --config 3 --mode shard --exchange inbox
run --config 4 --mode shard --exchange inbox
run --config 2 --mode shard --exchange inbox
run --config 2 --mode views
