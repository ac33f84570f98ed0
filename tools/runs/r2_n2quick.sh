set -u
O=gpurun_out/r2_n2q; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=30000000000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 bench.py --gpus 2 --steps 50 --warmup 10 --no-e2e --data-parallel-threshold 320000 > $O/a.log 2>&1; echo "dp forced rc=$?"; grep -E '^\{' $O/a.log | tail -1 | cut -c1-330
DE_B200_STREAM_PUSH=1 timeout 300 $TR --master-port 29702 bench.py --gpus 2 --steps 30 --warmup 10 --no-e2e --data-parallel-threshold 320000 > $O/b.log 2>&1; echo "dp + push rc=$?"; grep -E '^\{' $O/b.log | tail -1 | cut -c1-200
timeout 300 python -m pytest tests/test_dist_gpu.py -m gpu -q -x -p no:cacheprovider -k "replicated or dlrm_fast_world2" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
