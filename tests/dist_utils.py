"""Launch a multi-rank test case: one process per rank, rendezvous on 127.0.0.1."""
import os
import socket
import sys
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, case, device_type, backend, kwargs, errq):
  try:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(world)
    torch.set_num_threads(1)
    if device_type == "cuda":
      torch.cuda.set_device(rank)
      device = f"cuda:{rank}"
      dist.init_process_group("nccl", rank=rank, world_size=world,
                              device_id=torch.device(device))
    else:
      device = "cpu"
      dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import dist_cases
    getattr(dist_cases, case)(rank, world, device, backend, **kwargs)
    if device_type == "cuda":
      torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
  except Exception:  # pylint: disable=broad-except
    errq.put((rank, traceback.format_exc()))
    raise


def launch(case, world=2, device_type="cpu", backend="auto", timeout=240, **kwargs):
  ctx = mp.get_context("spawn")
  errq = ctx.SimpleQueue()
  port = _free_port()
  procs = []
  for r in range(world):
    p = ctx.Process(target=_worker, args=(r, world, port, case, device_type, backend, kwargs, errq))
    p.start()
    procs.append(p)
  failed = False
  for p in procs:
    p.join(timeout)
    if p.is_alive():
      p.terminate()
      failed = True
    elif p.exitcode != 0:
      failed = True
  msgs = []
  while not errq.empty():
    msgs.append(errq.get())
  if failed or msgs:
    for p in procs:
      if p.is_alive():
        p.terminate()
    detail = "\n".join(f"--- rank {r} ---\n{tb}" for r, tb in msgs) or "timeout / crash"
    raise AssertionError(f"distributed case {case} failed (world={world}):\n{detail}")
