#!/usr/bin/env python
"""Cross-rank timeline of one training step from the per-rank chrome traces written by
`bench.py --profile out.txt --profile-all-ranks` (out.txt.trace.json, out.txt.rankN.trace.json).

For the chosen step it prints, per rank, the kernels in launch order with start / duration relative
to the earliest kernel of the step, the idle gaps between them, and for every cross-GPU wait
(barrier / all-reduce kernels) which rank entered last - i.e. who the others were waiting for.

  python tools/critical_path.py gpurun_out/prof_n8.txt --step 2
"""
import argparse
import glob
import json
import re
from collections import defaultdict

# kernels that wait for peers at their head (flag spins folded into the data kernels)
WAIT_KERNELS = ("barrier_kernel", "allreduce_p2p_kernel", "allreduce_multimem_kernel",
                "push_segments_kernel", "lookup_fwd_kernel", "interact_fwd_kernel",
                "scatter_add_bwd_kernel", "scatter_add_staged_kernel", "segment_update_kernel",
                "stream_push_kernel", "sync_only_kernel")


def load_kernels(path):
  with open(path, encoding="utf-8") as f:
    trace = json.load(f)
  events = trace["traceEvents"] if isinstance(trace, dict) else trace
  ks = [e for e in events if e.get("ph") == "X" and e.get("cat", "").lower() in
        ("kernel", "gpu_memcpy", "gpu_memset")]
  ks.sort(key=lambda e: e["ts"])
  return ks


def short(name, n=44):
  name = re.sub(r"\(anonymous namespace\)::", "", name)
  name = re.sub(r"^void\s+", "", name)
  name = name.split("(")[0]
  return name[:n]


def split_steps(kernels, marker):
  """A step starts at every occurrence of the first kernel matching ``marker``."""
  idx = [i for i, k in enumerate(kernels) if marker in k["name"]]
  if not idx:
    return [kernels]
  # first occurrence per step: markers closer than half the median spacing belong to one step
  starts = [idx[0]]
  gaps = [kernels[b]["ts"] - kernels[a]["ts"] for a, b in zip(idx, idx[1:])]
  if gaps:
    med = sorted(gaps)[len(gaps) // 2]
    for a, b in zip(idx, idx[1:]):
      if kernels[b]["ts"] - kernels[a]["ts"] > 0.5 * med and b not in starts:
        starts.append(b)
  return [kernels[a:b] for a, b in zip(starts, starts[1:] + [len(kernels)])]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("profile", help="the --profile path given to bench.py")
  ap.add_argument("--step", type=int, default=2, help="which profiled step to print")
  ap.add_argument("--marker", default=None,
                  help="kernel that occurs once per step (splits the trace into steps); default: "
                       "push_segments_kernel if present (multi-GPU), else lookup_fwd_kernel")
  ap.add_argument("--gap-us", type=float, default=3.0, help="report idle gaps above this")
  args = ap.parse_args()

  paths = {0: args.profile + ".trace.json"}
  for p in glob.glob(args.profile + ".rank*.trace.json"):
    paths[int(re.search(r"\.rank(\d+)\.trace\.json$", p).group(1))] = p
  per_rank = {}
  for r, p in sorted(paths.items()):
    ks_all = load_kernels(p)
    marker = args.marker or ("push_segments_kernel" if any(
        "push_segments_kernel" in k["name"] for k in ks_all) else "lookup_fwd_kernel")
    steps = split_steps(ks_all, marker)
    per_rank[r] = steps[min(args.step, len(steps) - 1)]

  # ranks have their own clocks only if they are different hosts; one host -> one CUPTI clock
  t0 = min(ks[0]["ts"] for ks in per_rank.values() if ks)
  waits = defaultdict(list)
  for r, ks in per_rank.items():
    print(f"\n== rank {r}: {len(ks)} kernels, step span "
          f"{ks[-1]['ts'] + ks[-1]['dur'] - ks[0]['ts']:.1f} us, busy "
          f"{sum(k['dur'] for k in ks):.1f} us")
    prev_end = None
    occ = defaultdict(int)
    for k in ks:
      name = short(k["name"])
      if prev_end is not None and k["ts"] - prev_end > args.gap_us:
        print(f"      {'':44s}   idle {k['ts'] - prev_end:8.1f} us")
      print(f"  {k['ts'] - t0:9.1f} +{k['dur']:8.1f}  {name}")
      prev_end = max(prev_end or 0, k["ts"] + k["dur"])
      if any(w in k["name"] for w in WAIT_KERNELS):
        waits[(name, occ[name])].append((k["ts"] - t0, k["dur"], r))
        occ[name] += 1
  print("\n== cross-GPU waits: the rank that enters last is the one the others wait for")
  for (name, i), lst in sorted(waits.items(), key=lambda kv: min(x[0] for x in kv[1])):
    last = max(lst, key=lambda x: x[0])
    first = min(lst, key=lambda x: x[0])
    print(f"  {name}#{i}: first in rank {first[2]} @{first[0]:.1f}, last in rank {last[2]} "
          f"@{last[0]:.1f} (skew {last[0] - first[0]:.1f} us); durations "
          + " ".join(f"r{r}:{d:.0f}" for _, d, r in sorted(lst, key=lambda x: x[2])))


if __name__ == "__main__":
  main()
