"""CPU models of the two comparison networks the K2 epilogue relies on (csrc/fr_common.cuh: fold_sorted32 /
bitonic_merge_striped_desc; csrc/scan_mma.cu: reg_sort_desc / reg_select32_of64), stepped exactly as the device code
steps them -- lanes, slots, strides -- and checked against plain sorting.  The kernels themselves are covered by the
GPU parity tests; this pins the ALGORITHM (that the half-cleaner on the last block + one striped merge equals a fold,
and that the 32nd largest of 64 group maxima is a score 32 distinct rows reach) where no GPU is needed."""
import numpy as np
import pytest


def fold_sorted32_model(L, P, kpl):
    """L: 32*kpl keys sorted descending (0 = empty), entry i in lane i & 31, slot i >> 5; P: 32 keys sorted descending."""
    S = L.copy()
    for lane in range(32):  # L[KPL - 1] = max(L[KPL - 1], reverse32(p))
        S[(kpl - 1) * 32 + lane] = max(S[(kpl - 1) * 32 + lane], P[31 - lane])
    s = kpl // 2
    while s > 0:  # strides of 32 and more: slots of the same lane
        for j in range(kpl):
            if (j & s) == 0:
                for lane in range(32):
                    a, b = S[j * 32 + lane], S[(j + s) * 32 + lane]
                    S[j * 32 + lane], S[(j + s) * 32 + lane] = max(a, b), min(a, b)
        s //= 2
    s = 16
    while s > 0:  # strides below 32: one shuffle stage per slot
        T = S.copy()
        for j in range(kpl):
            for lane in range(32):
                y = S[j * 32 + (lane ^ s)]
                T[j * 32 + lane] = max(S[j * 32 + lane], y) if (lane & s) == 0 else min(S[j * 32 + lane], y)
        S = T
        s //= 2
    return S


@pytest.mark.parametrize("kpl", [1, 2, 4, 8])
def test_fold_is_one_half_cleaner_and_one_striped_merge(kpl):
    rng = np.random.default_rng(kpl)
    n = 32 * kpl
    for _ in range(200):
        n_list, n_pend = int(rng.integers(0, n + 1)), int(rng.integers(0, 33))
        vals = rng.permutation(100000)[: n_list + n_pend].astype(np.int64) + 1
        L = np.zeros(n, dtype=np.int64)
        L[:n_list] = np.sort(vals[:n_list])[::-1]
        P = np.zeros(32, dtype=np.int64)
        P[:n_pend] = np.sort(vals[n_list:])[::-1]
        want = np.sort(np.concatenate([L, P]))[::-1][:n]
        np.testing.assert_array_equal(fold_sorted32_model(L, P, kpl), want)


def reg_sort_desc_model(a):
    a = list(a)
    n = len(a)
    k = 2
    while k <= n:
        j = k >> 1
        while j > 0:
            for i in range(n):
                l = i ^ j
                if l > i:
                    hi, lo = max(a[i], a[l]), min(a[i], a[l])
                    a[i], a[l] = (hi, lo) if (i & k) == 0 else (lo, hi)
            j >>= 1
        k <<= 1
    return a


def select32_of64_model(g):
    a, b = reg_sort_desc_model(g[:32]), reg_sort_desc_model(g[32:])
    return min(max(a[i], b[31 - i]) for i in range(32))


@pytest.mark.parametrize("cols,group", [(256, 4), (128, 2)])
def test_first_tile_bound_is_reached_by_32_distinct_rows(cols, group):
    rng = np.random.default_rng(cols)
    for trial in range(200):
        if trial % 4 == 0:   # heavy ties
            scores = rng.integers(0, 5, size=cols).astype(np.float32) / 7.0
        else:
            scores = rng.standard_normal(cols).astype(np.float32)
        gmax = [float(scores[i * group:(i + 1) * group].max()) for i in range(64)]
        assert reg_sort_desc_model(gmax[:32]) == sorted(gmax[:32], reverse=True)
        bound = select32_of64_model(gmax)
        assert bound == sorted(gmax, reverse=True)[31]
        # the property the kernel needs: at least 32 distinct columns score at or above the bound, so it is a
        # lower bound on the tile's 32nd best score -- and every row of the true top 32 passes the gate "score >= bound"
        assert int((scores >= bound).sum()) >= 32
        assert bound <= np.sort(scores)[::-1][31]
