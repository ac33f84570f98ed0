set -u
O=gpurun_out/r2_final1; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=30000000000
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee $O/summary.txt; tail -1 $O/smoke.log | tee -a $O/summary.txt
timeout 100 python -m pytest tests/test_fused_optimizers.py tests/test_embedding_ops.py tests/test_embedding_layer.py tests/test_dlrm_fast.py -m gpu -q -x -p no:cacheprovider > $O/pytest_a.log 2>&1
echo "pytest a rc=$?" | tee -a $O/summary.txt; tail -2 $O/pytest_a.log | tee -a $O/summary.txt
timeout 170 python -m pytest tests/test_dist_gpu.py -m gpu -q -x -p no:cacheprovider -k "world1 and (multihot or ragged_dp or hybrid or row_slice or data_parallel or all_modes or synthetic_fast)" > $O/pytest_b.log 2>&1
echo "pytest b rc=$?" | tee -a $O/summary.txt; tail -2 $O/pytest_b.log | tee -a $O/summary.txt
