"""The schedule of the pipelined wm_host_step (wm_host_pipe_plan, wm_api.cu hp_schedule) -- host logic, no GPU.

A chunk of rows can be pushed as soon as it has arrived; the rim arrivals of tile row t (cell changers that cross a tile
boundary, at most one cell per step: common/particle.f90:83-169 under the CFL limit) can be appended once tile rows t-1, t, t+1
have been pushed; tile row t is final -- may go back to the host -- once t-1, t, t+1 have been placed; and the slab's first and
last tile rows exchange particles with the neighbour ranks (common/boundary_periodic.f90:99-248: periodic in y over the rank
ring), so everything that touches them waits for the ring exchange.  These are the conditions under which the pipelined call
gives the results of upload + wm_step + download; they are checked here for every slab height and chunk size."""
import pytest

TY = 8


def _check(nyl, rows):
    from wumingpic2d_b200.api import host_pipe_plan
    ops = host_pipe_plan(nyl, rows)
    nty = -(-nyl // TY)
    chunk = max(TY, rows // TY * TY)
    pushed, placed, down = set(), set(), []
    ring = False
    last_push_end = 0
    for kind, a, b in ops:
        if kind == "push":
            assert a == last_push_end and a < b <= nty, "chunks are pushed in upload order, without gaps"
            assert b - a <= chunk // TY
            assert not ring, "the ring exchange follows the last push"
            pushed |= set(range(a, b))
            last_push_end = b
        elif kind == "place":
            assert 0 <= a < b <= nty
            for t in range(a, b):
                assert t not in placed, "tile row %d placed twice" % t
                for nb in ((t - 1) % nty, t, (t + 1) % nty):
                    assert nb in pushed, "tile row %d placed before its neighbour %d was pushed" % (t, nb)
                if t in (0, nty - 1):
                    assert ring, "an edge tile row placed before the ring exchange"
                placed.add(t)
        elif kind == "down":
            assert 0 <= a < b <= nyl and b - a <= chunk
            for r in range(a, b):
                t = r // TY
                for nb in ((t - 1) % nty, t, (t + 1) % nty):
                    assert nb in placed, "row %d sent back before tile row %d was placed" % (r, nb)
                if t in (0, nty - 1):
                    assert ring
            down += list(range(a, b))
        else:
            assert kind == "ring" and not ring
            assert pushed == set(range(nty)), "the ring exchange needs the leavers of every tile row"
            ring = True
    assert ring and pushed == set(range(nty)) and placed == set(range(nty))
    assert sorted(down) == list(range(nyl)), "every row goes back exactly once"
    return ops


@pytest.mark.parametrize("rows", [8, 16, 24, 64])
def test_schedule_invariants_for_every_slab_height(rows):
    for nyl in list(range(1, 200)) + [512, 1000, 1024, 4096]:
        _check(nyl, rows)


def test_schedule_of_the_benchmark_slab():
    """512 rows in chunks of 16: 32 pushes; the downloads follow two tile rows behind, 3 chunks are left for the end"""
    ops = _check(512, 16)
    kinds = [k for k, _, _ in ops]
    assert kinds.count("push") == 32 and kinds.count("ring") == 1
    after_ring = kinds[kinds.index("ring"):]
    assert after_ring.count("down") <= 3
    # in the steady state a push is followed by the placement of the two tile rows before its last one and by one download chunk
    i = kinds.index("push", 10)
    assert kinds[i:i + 4] == ["push", "place", "down", "push"]


def test_rows_are_rounded_to_whole_tile_rows():
    from wumingpic2d_b200.api import host_pipe_plan
    assert host_pipe_plan(64, 20) == host_pipe_plan(64, 16)
    assert host_pipe_plan(64, 3) == host_pipe_plan(64, 8)
