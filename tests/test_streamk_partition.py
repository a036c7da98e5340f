"""Host-side model of the stream-K work partition used by the opt-in `gemv_sk_kernel` (bitblas_b200/csrc/bb_gemv.cu): the
T = (N/16) * (K/256) chunks, ordered row block by row block, are cut into Wtot equal contiguous ranges handed out in reverse
warp order; the range that holds the BEGINNING of a row block owns it, every other range touching the block parks the partial
sums of its FIRST segment.  These properties are what make the device-side fix-up deadlock-free and complete:

  * every chunk belongs to exactly one range;
  * a range parks at most once, and only its first segment;
  * every row block has exactly one owner, and the ranges an owner waits for have larger range indices, i.e. SMALLER warp ids
    (they were dispatched no later than the owner).
"""
import pytest


def range_begin(ri, T, Wtot):
    return ri * T // Wtot          # sk_range_begin


def segments(ri, T, Wtot, CPR):
    """(row_block, first_chunk, n_chunks, parks, closes) for every segment of range ri, in processing order."""
    t, t1 = range_begin(ri, T, Wtot), range_begin(ri + 1, T, Wtot)
    out = []
    while t < t1:
        rb, kc = divmod(t, CPR)
        n = min(t1 - t, CPR - kc)
        out.append((rb, t, n, kc != 0, kc + n == CPR))
        t += n
    return out


@pytest.mark.parametrize("N,K,Wtot", [(12288, 12288, 2368), (8192, 8192, 2368), (28672, 8192, 2368), (8192, 28672, 2368),
                                       (32, 1024, 8), (96, 2048, 8), (2080, 3072, 2368), (512, 8192, 1184), (4096, 256, 2368)])
def test_partition_invariants(N, K, Wtot):
    CPR = K // 256
    T = (N // 16) * CPR
    covered = [0] * T
    owners = {}
    parked = {}
    for ri in range(Wtot):
        segs = segments(ri, T, Wtot, CPR)
        for i, (rb, t, n, parks, closes) in enumerate(segs):
            for c in range(t, t + n):
                covered[c] += 1
            if parks:
                assert i == 0, "only the first segment of a range can start inside a row block"
                parked.setdefault(rb, []).append(ri)
            else:
                assert rb not in owners
                owners[rb] = (ri, closes)
    assert all(c == 1 for c in covered)
    assert set(owners) == set(range(N // 16))
    for rb, (ri, closes) in owners.items():
        contributors = parked.get(rb, [])
        if closes:
            assert not contributors
        # the owner's gather loop: rj = ri+1, ri+2, ... while the range begins before the end of the row block
        rb_end = (rb + 1) * CPR
        walked = []
        rj = ri + 1
        while not closes and rj < Wtot:
            b, e = range_begin(rj, T, Wtot), range_begin(rj + 1, T, Wtot)
            if b >= rb_end:
                break
            if b != e:
                walked.append(rj)
            if e >= rb_end:
                break
            rj += 1
        assert walked == sorted(contributors)
        # reverse mapping: warp id = Wtot - 1 - range index, so the owner only waits for warps with lower ids
        assert all((Wtot - 1 - rj) < (Wtot - 1 - ri) for rj in walked)
