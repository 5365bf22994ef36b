"""Host-side models of the work decompositions used by the opt-in kernels (same formulas as the device code), checked
exhaustively on small shapes: every unit is covered exactly once and the partial-result bookkeeping is consistent.
 - stream-K gradient kernel (grad_streamk.cuh: SkWalk, head / slot / n_other arithmetic; ctx.cu: upc, clusters, max_slots)
 - tile row pass (kernels.cu lse_tiles_kernel: (row block, slot) coverage and the per-block completion count)"""
import itertools

import pytest


def sk_plan(blocks, nkb, sms):
    total = blocks * nkb
    clusters = min(sms // 2, total)
    upc = (total + clusters - 1) // clusters
    clusters = (total + upc - 1) // upc
    max_slots = (nkb + upc - 1) // upc + 1
    return total, clusters, upc, max_slots


def sk_walk(cluster, upc, total, nkb):
    u, u1 = cluster * upc, min(total, (cluster + 1) * upc)
    while u < u1:
        block = u // nkb
        kb0 = u - block * nkb
        kb1 = min(nkb, kb0 + (u1 - u))
        u += kb1 - kb0
        yield block, kb0, kb1


@pytest.mark.parametrize("blocks,nkb,sms", [(64, 256, 148), (8, 256, 148), (1, 256, 148), (2, 7, 148), (5, 3, 8), (3, 256, 148),
                                            (64, 256, 132), (37, 31, 148), (200, 4, 148), (74, 256, 148), (75, 256, 148)])
def test_stream_k_bookkeeping(blocks, nkb, sms):
    total, clusters, upc, max_slots = sk_plan(blocks, nkb, sms)
    covered = {}
    heads, parts = {}, {}
    for w in range(clusters):
        segs = list(sk_walk(w, upc, total, nkb))
        assert segs, "no empty cluster"
        for idx, (b, kb0, kb1) in enumerate(segs):
            for kb in range(kb0, kb1):
                assert (b, kb) not in covered
                covered[(b, kb)] = w
            whole = kb0 == 0 and kb1 == nkb
            head = kb0 == 0 and not whole
            tile_u0 = b * nkb
            first_cluster = tile_u0 // upc
            last_cluster = min(clusters - 1, (tile_u0 + nkb - 1) // upc)
            if head:
                assert idx == len(segs) - 1, "a head is its cluster's last segment (so it never blocks the cluster's own work)"
                assert w == first_cluster
                heads[b] = last_cluster - w
                assert heads[b] >= 1
            elif not whole:
                slot = w - first_cluster - 1
                assert 0 <= slot < max_slots
                assert (b, slot) not in parts
                parts[(b, slot)] = (kb0, kb1)
    assert len(covered) == total
    for b, n_other in heads.items():
        got = sorted(s for (bb, s) in parts if bb == b)
        assert got == list(range(n_other)), (b, n_other, got)
        # slots are in K order
        ks = [parts[(b, s)][0] for s in got]
        assert ks == sorted(ks)
    # every partial belongs to a block that has a head
    assert {b for (b, _) in parts} == set(heads)


@pytest.mark.parametrize("N", [128, 129, 256, 1000, 8192 // 8, 1152])
def test_tile_row_pass_coverage(N):
    tb = (N + 127) // 128
    tiles = [(I, J) for I in range(tb - 1, -1, -1) for J in range(I, tb)]
    assert len(tiles) == tb * (tb + 1) // 2
    slots = {}
    cnt = [0] * tb
    for I, J in tiles:
        assert (I, J) not in slots                      # direct: rows of block I, slot J
        slots[(I, J)] = "direct"
        cnt[I] += 1
        if I < J:
            assert (J, I) not in slots                  # transposed: rows of block J, slot I
            slots[(J, I)] = "transposed"
            cnt[J] += 1
    assert all(c == tb for c in cnt)                    # the completion count the kernel waits for
    assert set(slots) == set(itertools.product(range(tb), range(tb)))
