set -u
O=gpurun_out/r2_n2q; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=30000000000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 bench.py --gpus 2 --steps 50 --warmup 10 > $O/a.log 2>&1; echo "default rc=$?"; grep -E '^\{' $O/a.log | tail -1 | cut -c1-330
timeout 300 $TR --master-port 29702 bench.py --gpus 2 --steps 50 --warmup 10 --no-e2e --data-parallel-threshold none > $O/b.log 2>&1; echo "dp none rc=$?"; grep -E '^\{' $O/b.log | tail -1 | cut -c1-200
DE_B200_STREAM_PUSH=0 timeout 300 $TR --master-port 29703 bench.py --gpus 2 --steps 50 --warmup 10 --no-e2e --no-verify > $O/c.log 2>&1; echo "pushoff rc=$?"; grep -E '^\{' $O/c.log | tail -1 | cut -c1-200
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 50 --warmup 10 > $O/d.log 2>&1; echo "n1 rc=$?"; grep -E '^\{' $O/d.log | tail -1 | cut -c1-200
tail -5 $O/a.log | cut -c1-300
