#!/usr/bin/env python
"""Summarise the source page of an ncu report: how many warp instructions every warp executes (straight line) and how
many are executed by only some warps / lanes (data-dependent paths), by opcode.
    ncu -i REPORT --page source --csv --kernel-name regex:NAME | python scripts/ncu_source_profile.py"""
import collections
import csv
import sys

rows = [r for r in csv.reader(sys.stdin) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
prof = []
for r in rows[1:]:
    try:
        prof.append((int(r[ix["Instructions Executed"]]), float(r[ix["Avg. Threads Executed"]]), r[ix["Source"]].strip(),
                     int(r[ix["# Samples"]])))
    except ValueError:
        continue
warps = prof[0][0]            # the entry instruction: executed once by every warp
total = sum(n for n, _, _, _ in prof)
samples = sum(s for _, _, _, s in prof)
print(f"warps {warps}, warp instructions {total} = {total / warps:.0f} per warp, SASS lines {len(prof)}")


def op(s):
    p = s.split()
    return (p[1] if p[0].startswith("@") else p[0]).split(".")[0]


for label, sel in (("executed once by every warp", lambda n: 0.98 * warps <= n <= 1.02 * warps),
                   ("loops (more executions than warps)", lambda n: n > 1.02 * warps),
                   ("data-dependent (fewer executions than warps)", lambda n: 0 < n < 0.98 * warps)):
    part = [(n, t, s, sm) for n, t, s, sm in prof if sel(n)]
    cnt = sum(n for n, _, _, _ in part)
    lanes = sum(n * t for n, t, _, _ in part) / max(cnt, 1)
    ops = collections.Counter()
    for n, _, s, _ in part:
        ops[op(s)] += n / warps
    print(f"{label}: {cnt / warps:.0f} per warp ({len(part)} lines), {lanes:.1f} lanes active on average, "
          f"{100.0 * sum(sm for _, _, _, sm in part) / max(samples, 1):.0f} % of the stall samples")
    print("   " + ", ".join(f"{k} {v:.0f}" for k, v in ops.most_common(18)))
