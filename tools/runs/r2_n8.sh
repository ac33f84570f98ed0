#!/bin/bash
# Round-2 8-GPU run: 1 -> 8 scaling of the headline on one box, A/B of the two overlap switches at
# N, per-rank graph-replay timeline at N.   gpurun --gpus 8 --timeout 1200 -- 'bash tools/runs/r2_n8.sh 8'
set -u
N=${1:-8}
O=gpurun_out/r2_n$N; mkdir -p $O
export DE_B200_FLAG_TIMEOUT_CYCLES=${DE_B200_FLAG_TIMEOUT_CYCLES:-30000000000}
run_bench() {  # n, tag, args...
  local n=$1 tag=$2; shift 2
  if [ $n -eq 1 ]; then
    CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --gpus 1 "$@" > $O/bench_${tag}_n1.log 2>&1
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > $O/bench_${tag}_n$n.log 2>&1
  fi
  local rc=$?
  grep -E '^\{' $O/bench_${tag}_n$n.log | tail -1 > $O/bench_${tag}_n$n.json
  python - "$O/bench_${tag}_n$n.json" "$tag n=$n rc=$rc" <<'PY' | tee -a $O/summary.txt
import json, sys
try:
  d = json.loads(open(sys.argv[1]).read())
  v = d.get("verify") or {}
  print("%-28s ms/step %.4f  %.2f M/s  e2e %s  verify %s (rel_l2 %s / %s)  loss %s" % (
      sys.argv[2], d.get("ms_per_step", -1), d.get("value", 0) / 1e6,
      ("%.2f M" % (d["e2e"]["value"] / 1e6)) if d.get("e2e") else None, v.get("ok"),
      v.get("rel_l2_err_of_update"), v.get("rel_l2_err_vs_autocast_oracle"), d.get("final_loss")))
except Exception as e:
  print(sys.argv[2], "no JSON:", e)
PY
}
run_bench $N fused --steps 50 --warmup 10
DE_B200_AR_OVERLAP=0 run_bench $N aroff --steps 50 --warmup 10 --no-e2e --no-verify
DE_B200_STREAM_PUSH=0 run_bench $N pushoff --steps 50 --warmup 10 --no-e2e --no-verify
DE_B200_STREAM_PUSH=0 DE_B200_AR_OVERLAP=0 run_bench $N bothoff --steps 50 --warmup 10 --no-e2e --no-verify
run_bench $N prof --steps 20 --warmup 5 --no-e2e --no-verify --profile $O/prof_n$N.txt \
  --profile-all-ranks --profile-graph 1
python tools/critical_path.py $O/prof_n$N.txt --step 2 > $O/critical_path_n$N.txt 2>&1
rm -f $O/prof_n$N.txt*.trace.json
for n in 4 2 1; do
  [ $n -lt $N ] || continue
  run_bench $n fused --steps 50 --warmup 10
done
