"""CPU mirrors of host-side index arithmetic of the CUDA library (no GPU, no oracle)."""
import random


def _stream_k_layout(ntiles, nblocks, nsm):
    """The balanced decomposition of glm_grad_kernel (glm_hmc.cu, QSB_FLAG_BALANCED): runs of U row-block units per CTA over the
    tile-major unit space; segment slot = CTA index - index of the first CTA touching the tile; the vector kernels read
    n_partials(tile) slots.  Returns {tile: [(slot, b0, b1), ...]}, U, the slot capacity the sampler allocates."""
    total = ntiles * nblocks
    U = max(12, (total + nsm - 1) // nsm)
    grid = (total + U - 1) // U
    capacity = (nblocks + U - 1) // U + 1
    segs = {}
    for c in range(grid):
        su, su_end = c * U, min((c + 1) * U, total)
        while su < su_end:
            tile = su // nblocks
            b0 = su - tile * nblocks
            b1 = min(nblocks, b0 + (su_end - su))
            segs.setdefault(tile, []).append((c - (tile * nblocks) // U, b0, b1))
            su += b1 - b0
    return segs, U, capacity


def test_stream_k_slots_cover_every_tile_exactly_once():
    """Every chain tile's row blocks are covered once, by contiguous segments, in slots 0..n_partials-1 (what the vector
    kernels sum), within the allocated capacity -- for the benchmark shapes and a few thousand random ones."""
    rnd = random.Random(1)
    shapes = [(t, nb, 148) for t in list(range(1, 40)) + [43, 86, 171, 683] for nb in (1, 2, 11, 12, 13, 26, 313, 314, 1000)]
    shapes += [(rnd.randint(1, 3000), rnd.randint(1, 4000), rnd.choice([148, 132, 108, 16])) for _ in range(2000)]
    for ntiles, nb, nsm in shapes:
        segs, U, capacity = _stream_k_layout(ntiles, nb, nsm)
        for t in range(ntiles):
            t0 = t * nb
            n_partials = ((t0 + nb - 1) // U - t0 // U) + 1          # n_partials_sk() in glm_hmc.cu
            s = sorted(segs[t])
            assert [x[0] for x in s] == list(range(n_partials)) and n_partials <= capacity, (ntiles, nb, nsm, t)
            assert s[0][1] == 0 and s[-1][2] == nb and all(s[i][2] == s[i + 1][1] and s[i][2] > s[i][1] for i in range(len(s) - 1))
