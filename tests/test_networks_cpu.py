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


def hist_count_model(hist, v):
    """mma_common.cuh: hist_count -- 16 coarse counters followed by 256 fine ones over [0, 1), scores below 1/16 skipped."""
    b = min(255, int(np.float32(v) * np.float32(256.0)))
    if b < 16:
        return
    hist[16 + b] += 1
    hist[b >> 4] += 1


def hist_edge_model(hist, ksel):
    """scan_mma.cu, the refresh: the lower edge of the highest bin with ksel counted rows at or above it (coarse bin
    first, then its 16 fine bins; fine counters that do not reach ksel leave the coarse edge); None = no bound yet."""
    cb, acc, above = -1, 0, 0
    for b in range(15, 0, -1):
        nacc = acc + hist[b]
        if cb < 0 and nacc >= ksel:
            cb, above = b, acc
        acc = nacc
    if cb < 1:
        return None
    fb, found = 0, False
    for j in range(15, 0, -1):
        above += hist[16 + 16 * cb + j]
        if not found and above >= ksel:
            fb, found = j, True
    return (16 * cb + fb) / 256.0


@pytest.mark.parametrize("ksel", [32, 128, 256])
def test_score_histogram_edge_is_a_score_ksel_counted_rows_reach(ksel):
    """The threshold K2 / K2s read off the shared histogram must never exceed the ksel-th best counted score (or a row
    of the final top-k' could be dropped), and it is the lower edge of that score's own bin when that score is above 1/8 -- also
    while the fine counters lag their coarse counter (a half-updated view), which may only loosen it."""
    rng = np.random.default_rng(ksel)
    for trial in range(60):
        n = int(rng.integers(ksel, 40 * ksel))
        sigma = [0.05, 0.15, 0.4][trial % 3]
        scores = np.clip(rng.standard_normal(n) * sigma + (0.0 if trial % 2 else 0.3), -1.0, 1.004).astype(np.float32)
        hist = [0] * (16 + 256)
        for v in scores:
            hist_count_model(hist, v)
        kth = float(np.sort(scores)[::-1][ksel - 1])
        edge = hist_edge_model(hist, ksel)
        if edge is None:
            assert int((scores >= 1.0 / 16.0).sum()) < ksel or kth < 2.0 / 16.0
            continue
        assert int((scores >= edge).sum()) >= ksel and edge <= kth
        if kth >= 2.0 / 16.0:  # with complete counters the edge is exactly the lower edge of the ksel-th best score's bin
            assert edge == min(255, int(np.float32(kth) * np.float32(256.0))) / 256.0
        # a lagging view: some fine increments have not landed yet -- the edge may only move down
        lag = list(hist)
        for b in rng.integers(16, 272, size=20):
            lag[b] = max(0, lag[b] - int(rng.integers(0, 3)))
        e2 = hist_edge_model(lag, ksel)
        assert e2 is None or e2 <= edge
