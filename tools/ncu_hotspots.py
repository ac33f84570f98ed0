#!/usr/bin/env python
"""Top stall sites of a kernel from an ``.ncu-rep`` (``ncu --set full --import-source on``):
the SASS instructions with the most warp-stall samples, with their share of all samples.

  python tools/ncu_hotspots.py gpurun_out/r2_n1/ncu_dlrm_n1.ncu-rep lookup_fwd_kernel [N]
"""
import csv
import io
import subprocess
import sys


def main():
  rep, kernel = sys.argv[1], sys.argv[2]
  top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
  out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"],
                       capture_output=True, text=True, check=True).stdout
  # the export holds one block per profiled launch: "Kernel Name" row, header row, SASS rows
  blocks, cur = [], None
  for row in csv.reader(io.StringIO(out)):
    if not row:
      continue
    if row[0] == "Kernel Name":
      cur = {"name": row[1], "header": None, "rows": []}
      blocks.append(cur)
    elif cur is not None and cur["header"] is None:
      cur["header"] = row
    elif cur is not None:
      cur["rows"].append(row)
  for b in blocks[:1]:  # first launch of the kernel in the report
    h = b["header"]
    i_src, i_all = h.index("Source"), h.index("# Samples")
    i_exec = h.index("Instructions Executed")
    rows = [(int(r[i_all] or 0), int(r[i_exec] or 0), r[i_src].strip(), k)
            for k, r in enumerate(b["rows"])]
    total = sum(r[0] for r in rows) or 1
    print(f"# {b['name'][:150]}")
    print(f"# {len(rows)} SASS instructions, {total} warp-stall samples; top {top} by samples")
    print(f"{'samples':>8} {'share':>6} {'executed':>10}  {'#':>5}  instruction")
    for s, e, src, k in sorted(rows, reverse=True)[:top]:
      print(f"{s:8d} {100.0 * s / total:5.1f}% {e:10d}  {k:5d}  {src}")


if __name__ == "__main__":
  main()
