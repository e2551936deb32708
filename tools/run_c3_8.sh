N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 60 --warmup 10 --config 3 --mode shard > /tmp/bench_out.txt 2>&1
grep -m3 "Error" /tmp/bench_out.txt; tail -1 /tmp/bench_out.txt > gpurun_out/c3_shard_$N.json; python -c "
import json,sys
j=json.load(open('gpurun_out/c3_shard_$N.json')); print(j['value'], j['unit'], j['ms_per_step'], 'e2e', j['e2e']['value']); print('  rank0', j['passes_us']); print('  max  ', j['passes_us_max_over_ranks'])
"
