python -m pytest tests -x -q -m gpu -k "fused_sharded" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 100 --warmup 10"
for cfg in 2 3 4; do
  for ex in inbox reduce; do
    echo "== config $cfg shard/$ex"; $TR --config $cfg --mode shard --exchange $ex 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
  j=json.loads(l); print(j['value'], j['unit'], j['ms_per_step'], j.get('passes_us'))
except Exception as e: print('ERR', l[-1500:])
"
  done
done
