#!/bin/bash
# Final 8-GPU DLRM measurements: NVLS vs P2P all-reduce, graph-mode kernel timeline, N=4.
mkdir -p gpurun_out; OUT=gpurun_out/final8.jsonl; : > $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
DE_B200_NVLS=1 timeout 240 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 100 --warmup 20 --profile gpurun_out/profile_n8_graph.txt --profile-graph 1 2>&1 | grep -E '^\{' | tail -1 >> $OUT
DE_B200_NVLS=0 timeout 240 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --steps 100 --warmup 20 --no-e2e 2>&1 | grep -E '^\{' | tail -1 >> $OUT
timeout 240 $TR --nproc-per-node 4 --master-port 29703 bench.py --gpus 4 --steps 100 --warmup 20 2>&1 | grep -E '^\{' | tail -1 >> $OUT
python -c "
import json
for l in open('$OUT'):
    d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],4), round(d['value']/1e6,2), (d.get('e2e') or {}).get('value'), d['config']['parallelism'][-40:])
"
