#!/usr/bin/env python
"""PCIe: H2D alone, D2H alone, both at once (pinned host memory, two streams) -- what a pipelined wm_host_step can reach."""
import json
import sys
import time

import torch

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gb * (1 << 30))
dev = torch.device("cuda", 0)
h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n, dtype=torch.uint8, device=dev)
d_dn = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, dn, chunk=None):
    torch.cuda.synchronize()
    t = time.perf_counter()
    ck = chunk or n
    for o in range(0, n, ck):
        if up:
            with torch.cuda.stream(s1):
                d_up[o:o + ck].copy_(h_up[o:o + ck], non_blocking=True)
        if dn:
            with torch.cuda.stream(s2):
                h_dn[o:o + ck].copy_(d_dn[o:o + ck], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t


out = {}
for name, a in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True)), ("both_12MB_chunks", (True, True, 12 << 20))):
    run(*a)
    ts = [run(*a) for _ in range(3)]
    out[name + "_GBps_per_direction"] = n / min(ts) / 1e9
print(json.dumps(out))
