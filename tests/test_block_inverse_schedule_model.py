"""CPU model of the progressive block-inverse schedule (`trtri_lower_step`, gpirt_b200/csrc/linalg.cu).

On the substitution route the backward pass  L^T A = X  (reference src/draw-fstar.cpp:7,24: solve(trimatu(L.t()), .))
runs on the inverses of max_block x max_block diagonal blocks of L.  They are built from the 128-block inverses that the
factorisation produces anyway, by recursive doubling  inv([[L11,0],[L21,L22]]) = [[X11,0],[-X22 L21 X11, X22]] — and, so
that almost nothing is left to do after the last panel, the merges are issued BEHIND the panels: step k (panel k is final)
copies the 128-block inverse of panel k and performs every merge that panel k completes (binary-counter pattern); the last
step also merges the ragged pairs.  This file restates that schedule in numpy, with the same conditions as the C++, and
checks that it (a) only ever reads rows / columns of L that are final at step k and (b) ends with the exact block inverses
for full and ragged orders.  The CUDA implementation is compared with the oracle in tests/test_gpu_parity.py
(test_chain_blocked_substitution_solves_match_inverse_route, test_graph_replayed_sweeps_equal_eager_sweeps)."""
import numpy as np
import pytest

NB = 128


def merge(L, X, o, s, rows, final):
    """X21 = -X22 (L21 X11) for the pair starting at o: first block order s, second block order rows"""
    assert o + s + rows <= final, "merge reads rows of L that are not final yet"
    L21 = L[o + s:o + s + rows, o:o + s]
    X11 = X[o:o + s, o:o + s]
    X22 = X[o + s:o + s + rows, o + s:o + s + rows]
    X[o + s:o + s + rows, o:o + s] = -X22 @ (L21 @ X11)


def step(L, X, n, max_block, k):
    """gpirt::trtri_lower_step: same control flow"""
    nblk = -(-n // NB)
    k0 = k * NB
    nb = min(NB, n - k0)
    X[k0:k0 + nb, k0:k0 + nb] = np.linalg.inv(L[k0:k0 + nb, k0:k0 + nb])   # what k_diag128 leaves in Dinv
    stop = min(n, max_block) if max_block > 0 else n
    done = k0 + nb
    s = NB
    merges = 0
    while s < stop:
        if done % (2 * s) == 0:
            merge(L, X, done - 2 * s, s, s, done)
            merges += 1
            s *= 2
            continue
        if k != nblk - 1:
            break
        o_r = (n // (2 * s)) * 2 * s
        s2 = min(s, n - (o_r + s))
        if s2 > 0:
            merge(L, X, o_r, s, s2, done)
            merges += 1
        s *= 2
    return merges


@pytest.mark.parametrize("n", [100, 128, 256, 300, 640, 1024, 1100, 1152, 2304, 2500])
@pytest.mark.parametrize("max_block", [256, 512, 1024])
def test_progressive_schedule_builds_the_block_inverses(n, max_block):
    rs = np.random.RandomState(n + max_block)
    L = np.tril(rs.randn(n, n)) * 0.05 + np.eye(n)
    X = np.zeros((n, n))
    nblk = -(-n // NB)
    per_step = [step(L, X, n, max_block, k) for k in range(nblk)]
    for r0 in range(0, n, max_block):
        r1 = min(n, r0 + max_block)
        want = np.linalg.inv(L[r0:r1, r0:r1])
        got = X[r0:r1, r0:r1]
        assert np.max(np.abs(got - want)) <= 1e-9 * max(1.0, np.max(np.abs(want)))
    # nothing outside the diagonal blocks is ever written (the caller zeroes X once)
    mask = np.zeros((n, n), dtype=bool)
    for r0 in range(0, n, max_block):
        mask[r0:r0 + max_block, r0:r0 + max_block] = True
    assert np.all(X[~mask] == 0.0)
    # the point of the schedule: after the last panel at most one merge per level is left
    levels = int(np.log2(max_block // NB)) if max_block > NB else 0
    assert per_step[-1] <= levels


def test_full_order_leaves_exactly_one_merge_per_level_in_the_tail():
    n, max_block = 4096, 1024
    L = np.eye(n)
    X = np.zeros((n, n))
    per_step = [step(L, X, n, max_block, k) for k in range(n // NB)]
    assert per_step[-1] == 3                      # 128 -> 256 -> 512 -> 1024 of the last block
    assert sum(per_step) == (n // NB) - n // max_block   # a binary tree per 1024-block: 7 merges each
