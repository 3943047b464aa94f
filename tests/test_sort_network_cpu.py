"""The stage schedule of ``block_sort`` (csrc/raster_fwd.cu) replayed on the CPU: the per-tile bitonic network in which
every warp finishes the stages that stay inside its own chunk of the list behind warp barriers only.  Checks, for list
lengths around every power of two up to beyond both shared-memory capacities, that (1) the warp-local phases touch only
the warp's chunk — so running the warps in any order is legitimate — and (2) the result is sorted.  The GPU tests
compare the kernel's sorted tile lists with the oracle's bit for bit; this one pins the index arithmetic without a GPU."""
import random

import pytest

K_THREADS, WARPS = 256, 8


def _cmpx(k, i, j):
    if k[i] > k[j]:
        k[i], k[j] = k[j], k[i]


def _flip(k, n, lsize, t_begin, t_end, t_step, touched=None):
    size, lhs = 1 << lsize, lsize - 1
    hs = size >> 1
    for t in range(t_begin, t_end, t_step):
        blk, w = t >> lhs, t & (hs - 1)
        i, j = (blk << lsize) + w, (blk << lsize) + (size - 1 - w)
        if touched is not None:
            touched.update((i, j))
        if j < n:
            _cmpx(k, i, j)


def _disperse(k, n, lstep, t_begin, t_end, t_step, touched=None):
    step = 1 << lstep
    for t in range(t_begin, t_end, t_step):
        i = ((t >> lstep) << (lstep + 1)) + (t & (step - 1))
        j = i + step
        if touched is not None:
            touched.update((i, j))
        if j < n:
            _cmpx(k, i, j)


def block_sort(k, n, rng):
    if n < 2:
        return 0
    lpad = 1
    while (1 << lpad) < n:
        lpad += 1
    half = 1 << (lpad - 1)
    if lpad <= 6:
        for lsize in range(1, lpad + 1):
            for lane in range(32):
                _flip(k, n, lsize, lane, half, 32)
            for lstep in range(lsize - 2, -1, -1):
                for lane in range(32):
                    _disperse(k, n, lstep, lane, half, 32)
        return 1
    lchunk = lpad - 3
    cp, chunk, barriers = 1 << (lchunk - 1), 1 << lchunk, 0

    def local(ops):
        for warp in rng.sample(range(WARPS), WARPS):          # a whole warp-local phase at a time, warps in random order
            touched = set()
            for kind, l in ops:
                for lane in range(32):
                    (_flip if kind == "f" else _disperse)(k, n, l, warp * cp + lane, (warp + 1) * cp, 32, touched)
            assert all(warp * chunk <= e < (warp + 1) * chunk for e in touched)

    pending = []
    for lsize in range(1, lpad + 1):
        if lsize <= lchunk:
            pending += [("f", lsize)] + [("d", l) for l in range(lsize - 2, -1, -1)]
            if lsize == lchunk:
                local(pending)
                pending, barriers = [], barriers + 1
        else:
            for tid in range(K_THREADS):
                _flip(k, n, lsize, tid, half, K_THREADS)
            barriers += 1
            lstep = lsize - 2
            while lstep >= lchunk:
                for tid in range(K_THREADS):
                    _disperse(k, n, lstep, tid, half, K_THREADS)
                barriers, lstep = barriers + 1, lstep - 1
            local([("d", l) for l in range(lstep, -1, -1)])
            barriers += 1
    return barriers


@pytest.mark.parametrize("n", list(range(0, 70)) + [127, 128, 129, 255, 256, 257, 1000, 1025, 2047, 2048, 2049, 4096, 4097, 4989, 8192, 9001])
def test_block_sort_schedule_sorts_and_keeps_warps_independent(n):
    rng = random.Random(n)
    for hi in (40, 1 << 50):                                   # many ties / 64-bit keys
        k = [rng.randrange(hi) for _ in range(n)]
        ref = sorted(k)
        barriers = block_sort(k, n, rng)
        assert k == ref
    if n == 2048:
        assert barriers == 10                                  # block barriers; the one-barrier-per-stage network needs 66
