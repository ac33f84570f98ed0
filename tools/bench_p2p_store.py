#!/usr/bin/env python
"""What does it take to saturate NVLink with kernel-issued stores?  (2 GPUs)

  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_p2p_store.py

Rank 0 stores rows of a local buffer into rank 1's memory (and into its own, for reference) with
the access shapes the exchange kernels use: bytes per lane, contiguous bytes per row, rows
scattered with a larger destination stride, grid size.  Both directions run at once in the
"bidir" rows (what the training step does).  Prints JSON lines (GB/s)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from distributed_embeddings_b200.ops import _native
from distributed_embeddings_b200.parallel.comm import CommContext


def main():
  rank = int(os.environ.get("RANK", "0"))
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
  ops = _native.require()
  ctx = CommContext.default(torch.device("cuda", rank))
  total = 256 << 20
  buf = ctx.alloc(2 * total, "p2p_bench")  # destination rows may be strided 2x
  src = torch.randint(0, 255, (total,), dtype=torch.uint8, device="cuda")
  peer = buf.ptrs[1 - rank]
  local = buf.ptrs[rank]
  sms = torch.cuda.get_device_properties(rank).multi_processor_count
  rows_out = []

  def run(dst, row_bytes, vec, stride_mul, unroll, bps, threads, active):
    n_rows = total // row_bytes
    dst_stride = row_bytes * stride_mul
    fn = lambda: ops.p2p_store_bench(src, dst, n_rows, row_bytes, vec, dst_stride, unroll,
                                     sms * bps, threads)
    dist.barrier()
    torch.cuda.synchronize()
    if not active:
      dist.barrier()
      return None
    for _ in range(2):
      fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
      fn()
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    return total * 5 / (a.elapsed_time(b) * 1e-3) / 1e9

  configs = []
  for row_bytes in (128, 256, 512, 1024):
    for vec in (4, 8, 16):
      if row_bytes // vec > 32 and (row_bytes // vec) % 32:
        continue
      for stride_mul in (1, 2):
        configs.append((row_bytes, vec, stride_mul, 4, 4, 256))
  for bps, threads in ((1, 256), (2, 256), (8, 256), (2, 1024), (1, 512)):
    configs.append((256, 16, 2, 4, bps, threads))
    configs.append((256, 8, 2, 4, bps, threads))
  for unroll in (1, 2, 8):
    configs.append((256, 16, 2, unroll, 4, 256))
  for cfg in configs:
    row_bytes, vec, stride_mul, unroll, bps, threads = cfg
    uni = run(peer, *cfg, active=rank == 0)           # one direction
    bid = run(peer, *cfg, active=True)                # both directions at once
    loc = run(local, *cfg, active=rank == 0)          # same kernel, local memory
    if rank == 0:
      rec = {"row_bytes": row_bytes, "vec_bytes": vec, "dst_stride_x": stride_mul,
             "unroll": unroll, "blocks_per_sm": bps, "threads": threads,
             "peer_unidir_GBps": round(uni, 1), "peer_bidir_GBps": round(bid, 1),
             "local_GBps": round(loc, 1)}
      rows_out.append(rec)
      print(json.dumps(rec), flush=True)
  if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows_out, open("gpurun_out/bench_p2p_store.json", "w"), indent=1)
  dist.destroy_process_group()


if __name__ == "__main__":
  main()
