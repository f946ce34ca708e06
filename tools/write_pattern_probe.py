#!/usr/bin/env python3
"""How fast can this box WRITE the two H planes of one ML-20M-sized cluster (26 744 x 27 136 fp64 + 4-byte) when the
bursts in flight are scattered the way a (row, column range) task grid scatters them?  torch only (plumbing): each
"task" writes an 8 KB fp64 burst + a 4 KB burst; task order = range-major (rows 214 KB apart in flight) or row-major."""
import json, sys, torch
I, ld, RW = 26744, 27136, 1024
dev = torch.device("cuda")
H = torch.empty((I, ld), dtype=torch.float64, device=dev)
Hh = torch.empty((I, ld), dtype=torch.float32, device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {"bytes": H.numel() * 12}
out["fill_contiguous_ms"] = t(lambda: (H.fill_(1.0), Hh.fill_(1.0)))
# column-block at a time = what a range-major grid has in flight: rows 214 KB apart
def colblocks():
    for c in range(0, ld, RW):
        H[:, c:c + RW].fill_(1.0); Hh[:, c:c + RW].fill_(1.0)
out["fill_column_blocks_ms"] = t(colblocks, 2)
for k in ("fill_contiguous_ms", "fill_column_blocks_ms"):
    out[k.replace("_ms", "_TBps")] = out["bytes"] / out[k] / 1e9
print(json.dumps(out))
