"""Host logic of the persistent CG kernel (cg_persist_kernel.cu): the block decomposition of a slab over the SMs."""
import pytest


@pytest.mark.parametrize("nx,nyl", [(4096, 512), (4096, 513), (256, 64), (2048, 512), (16384, 128), (1024, 1024), (56, 28),
                                    (40, 19), (8, 2), (5, 3), (208, 150), (4097, 511)])
def test_plan_covers_the_slab(nx, nyl):
    """Every BASELINE slab (config 1: 256 x 64 per rank; 2: 4096 x 512; 3: 2048 x 512; 4: 16384 x 128) fits: at most 148 CTAs,
    a thread owns a vertical run of at most CGP_K = 16 cells, columns x row groups <= 1024 threads, the column-major tile
    (odd column pitch) fits 225 KB of shared memory, blocks at least 4 cells wide (wall rule)."""
    import wumingpic2d_b200.api as A
    cbx, cby, smem, rl = A.cg_plan(nx, nyl)
    assert 1 <= cbx * cby <= 148
    bw, bh = -(-nx // cbx), -(-nyl // cby)
    assert 1 <= rl <= 16
    assert bw * -(-bh // rl) <= 1024
    assert (bw + 2) * ((bh + 2) | 1) * 8 == smem <= 227 * 1024 - 2048
    assert nx // cbx >= 4 or cbx == 1
    assert nyl // cby >= 1


def test_plan_rejects_what_does_not_fit():
    import wumingpic2d_b200.api as A
    with pytest.raises(A.WmError):
        A.cg_plan(2048, 2048)       # 4.2 M cells > 148 x 1024 x 16: the host-loop CG takes over
    with pytest.raises(A.WmError):
        A.cg_plan(4096, 512, nsm=64)
