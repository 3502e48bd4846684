"""CPU worker of tests/test_gloo.py (torch.distributed, backend gloo, world_size 2): the host-side logic of the N > 1
path that needs no GPU -- slab bounds per rank (common/mpi_set.f90:36-47 remainder rule) as bench.py / mgpu_worker.py
derive them, the ring neighbours, the id broadcast that precedes wm_comm_init, and the max / sum reductions bench.py
reports with."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    # ---- slab bounds: every rank derives all of them from the same oracle world
    prm = O.weibel_params(24, 8 * world + 3, 4, nranks=world)   # ny not divisible by the rank count
    w = O.World(prm)
    nys, nye = w.bounds(rank)
    mine = torch.tensor([nys, nye], dtype=torch.int64)
    allb = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allb, mine)
    rows = [int(b[1] - b[0] + 1) for b in allb]
    assert sum(rows) == prm["ny"] and max(rows) - min(rows) <= 1 and rows == sorted(rows, reverse=True)
    assert int(allb[0][0]) == prm["nygs"] and all(int(allb[k + 1][0]) == int(allb[k][1]) + 1 for k in range(world - 1))
    nup, ndown = (rank + 1) % world, (rank - 1) % world         # mpi_set.f90:44-47
    # ---- ring exchange of one row vector (stand-in for the ghost rows): what I get from ndown is its last row
    row = torch.full((8,), float(nye))
    got = torch.zeros(8)
    reqs = [dist.isend(row, nup), dist.irecv(got, ndown)]
    for r in reqs:
        r.wait()
    assert float(got[0]) == float(allb[ndown][1])
    # ---- unique-id broadcast (bench.py / mgpu_worker.py do this before wm_comm_init)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    # ---- reductions of the bench harness: time = max over ranks, particles = sum over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(rows[rank] * prm["nx"])], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    assert float(t) == float(world) and float(n) == prm["nx"] * prm["ny"]
    w.close()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d/%d ok" % (rank, world))


if __name__ == "__main__":
    main()
