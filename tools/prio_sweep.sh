#!/bin/bash
# stream-priority / occupancy-cap sweep on one GPU, config given as $1
cfg=$1; shift
run() { VCT_MAIN_STREAM_PRIORITY=$1 python bench.py --config $cfg --steps 40 --warmup 5 --no-cpu-baseline --no-strong --tune "$2" 2>/dev/null | python -c "
import sys,json; t=sys.stdin.read(); d=json.loads(t[t.index('{'):]); print('main_prio=$1 tune=$2 ->', d['value'], 'fps  e2e', d['e2e']['value'])"; }
run 0 ""
run -1 "SideStreamsLowPriority=1"
run -1 "SideStreamsLowPriority=1,ConeSmemPad=24000"
run -1 "SideStreamsLowPriority=1,ConeSmemPad=33000"
run -1 "SideStreamsLowPriority=1,ConeSmemPad=33000,ChainBlockThreads=128"
run -1 "SideStreamsLowPriority=1,ConeSmemPad=33000,ChainBlockThreads=64,RasterBlockThreads=64"
run -1 "SideStreamsLowPriority=1,ConeSmemPad=24000,ChainBlockThreads=64,RasterBlockThreads=64"
run 0 "ConeSmemPad=33000,ChainBlockThreads=64,RasterBlockThreads=64"
