"""CPU: a numpy restatement of what the subsample tensor-core kernel does between the per-locus selection and the 256
counts (tetrad_b200/csrc/tq_resolve.cu: select_chunk_codes, the register stack of code planes, scatter_overflow, the tile
product) -- checked against the direct histogram of the selected loci.  It pins the ALGORITHM (every selected locus is
counted exactly once, by a product or by the walk, whatever the stacking limit); the CUDA code itself is held to the
oracle by the `-m gpu` tests (test_stacked_tiles_count_every_locus_once runs the same limits on the device)."""
import numpy as np
import pytest


def select_tile(codes, valid):
    """codes u8[K, 4, 1024] (0..3), valid bool[K, 4, 1024]: rank chunks of one tile for the four tips of a quartet.
    Returns (taken bool[1024], sel_code u8[4, 1024]): the first rank whose column is complete and segregating."""
    taken = np.zeros(1024, bool)
    sel_code = np.zeros((4, 1024), np.uint8)
    for k in range(codes.shape[0]):
        complete = valid[k].all(0)
        differs = (codes[k] != codes[k][0]).any(0)
        sel = complete & differs & ~taken
        taken |= sel
        sel_code[:, sel] = codes[k][:, sel]
    return taken, sel_code


def product(occ, stack_code):
    """The 16 x 16 x 1024 tile product on one-hot planes expanded from the stack's code planes."""
    T = np.zeros(256, np.int64)
    onehot = [[(occ & (stack_code[t] == v)) for v in range(4)] for t in range(4)]
    for a in range(4):
        for b in range(4):
            x = onehot[0][a] & onehot[1][b]
            for c in range(4):
                for d in range(4):
                    T[64 * a + 16 * b + 4 * c + d] += int((x & onehot[2][c] & onehot[3][d]).sum())
    return T


def count_quartet(tiles, stack_ovf):
    """tiles: list of (codes, valid).  Lanes own 32 consecutive slots; a tile is stacked iff no lane has more than
    `stack_ovf` selected loci whose slot is taken (those are walked one by one)."""
    T = np.zeros(256, np.int64)
    occ = np.zeros(1024, bool)
    stack_code = np.zeros((4, 1024), np.uint8)
    stacked = False
    nproducts = nwalked = 0
    for item in tiles + [None]:
        last = item is None
        taken, sel_code = (np.zeros(1024, bool), np.zeros((4, 1024), np.uint8)) if last else select_tile(*item)
        ovf = taken & occ
        worst = ovf.reshape(32, 32).sum(1).max()
        if stacked and (last or worst > stack_ovf):
            T += product(occ, stack_code)
            nproducts += 1
            stacked = False
        if last:
            break
        if not stacked:
            occ, stack_code, stacked = taken.copy(), sel_code.copy(), True
        else:
            put = taken & ~occ
            stack_code[:, put] = sel_code[:, put]
            occ |= taken
            for s in np.flatnonzero(ovf):                     # the walk: one atomic per locus
                a, b, c, d = sel_code[:, s]
                T[64 * a + 16 * b + 4 * c + d] += 1
                nwalked += 1
    return T, nproducts, nwalked


@pytest.mark.parametrize("stack_ovf", [0, 3, 8, 32])
def test_stacking_counts_every_selected_locus_once(stack_ovf):
    rng = np.random.default_rng(5)
    tiles = []
    want = np.zeros(256, np.int64)
    for t in range(9):
        K = int(rng.integers(1, 6))
        codes = rng.integers(0, 4, size=(K, 4, 1024)).astype(np.uint8)
        # few usable columns, as in the data: most columns invariant, a tenth missing per tip
        invariant = rng.random((K, 1024)) < (0.55 + 0.04 * t)
        codes[:, 1:, :] = np.where(invariant[:, None, :], codes[:, :1, :], codes[:, 1:, :])
        valid = rng.random((K, 4, 1024)) > 0.12
        tiles.append((codes, valid))
        taken, sel_code = select_tile(codes, valid)
        idx = 64 * sel_code[0].astype(int) + 16 * sel_code[1] + 4 * sel_code[2] + sel_code[3]
        np.add.at(want, idx[taken], 1)
    got, nproducts, nwalked = count_quartet(tiles, stack_ovf)
    assert (got == want).all()
    assert not got[[0, 85, 170, 255]].any()                   # a selected column is never invariant
    if stack_ovf == 0:
        assert nwalked == 0
    if stack_ovf == 32:
        assert nproducts == 1 and nwalked > 0                 # everything on one stack, the rest walked
