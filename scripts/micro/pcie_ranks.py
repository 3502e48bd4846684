#!/usr/bin/env python
"""All ranks of a node move pinned host memory over PCIe at once: what bounds wm_host_step at N = 4, 8?

    python -m torch.distributed.run --nproc-per-node N scripts/micro/pcie_ranks.py [GB per buffer]

Per rank: H2D alone, D2H alone, both at once, for three kinds of host buffer -- (a) cudaHostAlloc (torch pin_memory) made BEFORE
the rank is bound to the CPUs next to its GPU, (b) the same made AFTER the binding (what bench.py does), (c) ordinary pages
first-touched after the binding and page-locked with cudaHostRegister (wm_host_register).  Rank 0 prints one JSON line with
the per-rank rates and their sums.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import bind_to_gpu_numa  # noqa: E402


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    n = int(gb * (1 << 30))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    d_up = torch.empty(n, dtype=torch.uint8, device=dev)
    d_dn = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(h_up, h_dn, up, dn):
        ck = 12 << 20
        barrier()
        t = time.perf_counter()
        for o in range(0, n, ck):
            if up:
                with torch.cuda.stream(s1):
                    d_up[o:o + ck].copy_(h_up[o:o + ck], non_blocking=True)
            if dn:
                with torch.cuda.stream(s2):
                    h_dn[o:o + ck].copy_(d_dn[o:o + ck], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        barrier()
        return n / dt / 1e9

    def measure(h_up, h_dn):
        out = {}
        for name, a in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
            run(h_up, h_dn, *a)
            out[name] = max(run(h_up, h_dn, *a) for _ in range(2))
        return out

    res = {}
    a_up, a_dn = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    res["hostalloc_before_binding"] = measure(a_up, a_dn)
    del a_up, a_dn
    cpus = bind_to_gpu_numa(local)
    b_up, b_dn = torch.empty(n + 4096, dtype=torch.uint8).pin_memory()[:n], torch.empty(n + 4096, dtype=torch.uint8).pin_memory()[:n]
    res["hostalloc_after_binding"] = measure(b_up, b_dn)
    del b_up, b_dn
    import wumingpic2d_b200.api as api
    c_up, c_dn = np.ones(n, dtype=np.uint8), np.ones(n, dtype=np.uint8)       # first touch here, after the binding
    api.host_register(c_up)
    api.host_register(c_dn)
    res["registered_after_binding"] = measure(torch.from_numpy(c_up), torch.from_numpy(c_dn))
    api.host_unregister(c_up)
    api.host_unregister(c_dn)
    node = None
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
    except Exception:
        bus = None
    mine = {"rank": rank, "cpus": cpus, "rates_GBps": res}
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        tot = {k: {m: sum(r["rates_GBps"][k][m] for r in allr) for m in ("h2d", "d2h", "both")} for k in res}
        numa = ""
        try:
            numa = open("/sys/devices/system/node/online").read().strip()
        except Exception:
            pass
        print(json.dumps({"n_gpus": world, "GB_per_buffer": gb, "numa_nodes_online": numa, "host_cpus": os.cpu_count(),
                          "sum_GBps_per_direction": tot, "ranks": allr}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
