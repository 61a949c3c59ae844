"""Executable model of the warp-per-pixel exact median of csrc/collapse.cu (``collapse_median_warp_kernel``): the same
integer/key logic in numpy, checked against ``np.nanmedian`` bit for bit on random and adversarial columns, with the
pass counts the kernel's cost model is built on.  ``python tools/median_bracket_model.py``"""
import numpy as np

PAD = np.uint32(0xFFFFFFFF)


def f2key(v):
    b = np.float32(v).view(np.uint32)
    return np.uint32(~b) if b & np.uint32(0x80000000) else np.uint32(b | np.uint32(0x80000000))


def key2f(k):
    k = np.uint32(k)
    b = np.uint32(k & np.uint32(0x7FFFFFFF)) if k & np.uint32(0x80000000) else np.uint32(~k)
    return b.view(np.float32)


def median_model(col, stats=None):
    n = len(col)
    keys = np.array([PAD if v != v else f2key(v) for v in col], dtype=np.uint32)
    m = int(np.sum(keys != PAD))
    if m == 0:
        return np.float32(np.nan)
    r1, r2 = (m - 1) >> 1, m >> 1
    s = np.sort(keys[:32])
    ns = int(np.sum(s != PAD))
    lo, hi, clo, chi = 0, 0xFFFFFFFF, 0, m
    have_lo = have_hi = False
    passes = 0
    it = 0
    k1 = k2 = None
    if ns > 0:
        idx = ns >> 1
        p = int(s[idx])
    else:
        idx = 0
        p = lo + ((hi - lo) >> 1)
    first = True
    while True:
        c = int(np.sum(keys < np.uint32(p)))
        passes += 1
        if first and ns > 1 and (s[idx] == s[min(idx + 1, ns - 1)] or s[idx] == s[max(idx - 1, 0)]):
            cle = int(np.sum(keys <= np.uint32(p)))
            passes += 1
            if c <= r1 and r2 < cle:
                k1 = k2 = p
                break
        first = False
        if c <= r1:
            lo, clo, have_lo = p, c, True
        elif c > r2:
            hi, chi, have_hi = p, c, True
        else:                                    # c == r2 == r1 + 1: the pivot splits the two middle ranks
            k1 = int(keys[keys < np.uint32(p)].max())
            k2 = int(keys[keys >= np.uint32(p)].min())
            passes += 1
            break
        if chi - clo <= 32:
            cand = np.sort(keys[(keys >= np.uint32(lo)) & (keys < np.uint32(hi))])
            assert len(cand) == chi - clo
            k1, k2 = int(cand[r1 - clo]), int(cand[r2 - clo])
            break
        if hi - lo <= 1:
            k1 = k2 = lo
            break
        it += 1
        use_mid = it >= 6 and it % 3 == 0
        pn = None
        if not use_mid and not (have_lo and have_hi) and ns > 0:
            d = ((r1 - c) * ns) // m if c <= r1 else -(((c - r1) * ns) // m)
            g = min(it, 3)
            step = d + g if c <= r1 else d - g
            idx = min(max(idx + step, 0), ns - 1)
            pn = int(s[idx])
        elif not use_mid and have_lo and have_hi:
            flo, fhi = np.float32(key2f(lo)), np.float32(key2f(hi))
            t = np.float32((r1 + 0.5 - clo) / (chi - clo))
            t = min(max(t, np.float32(0.15)), np.float32(0.85))
            with np.errstate(all="ignore"):
                pf = np.float32(flo + np.float32(fhi - flo) * t)
            pn = int(f2key(pf)) if pf == pf else None
        if pn is not None and lo < pn < hi:
            p = pn
        else:
            p = lo + ((hi - lo) >> 1)
    if stats is not None:
        stats.append(passes)
    a, b = key2f(k1), key2f(k2)
    if m & 1:
        return a
    return np.float32((a + b) * np.float32(0.5))


def _check(col, stats):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.nanmedian(col.astype(np.float32))
    got = median_model(col.astype(np.float32), stats)
    assert (got != got and want != want) or got.view(np.uint32) == np.float32(want).view(np.uint32) or \
        (got == want), (got, want)           # -0.0 == 0.0: np.nanmedian of ties at zero may return either sign


def main():
    rng = np.random.default_rng(0)
    for name, gen, cnt in (
            ("gauss 500", lambda: rng.normal(size=500), 1500),
            ("gauss 499", lambda: rng.normal(size=499), 500),
            ("student-t2 500", lambda: rng.standard_t(2, size=500), 500),
            ("lognormal 1000", lambda: rng.lognormal(size=1000), 300),
            ("5 huge outliers", lambda: rng.permutation(np.concatenate([rng.normal(size=495), 1e30 * rng.normal(size=5)])), 300),
            ("30% NaN", lambda: np.where(rng.random(500) < 0.3, np.nan, rng.normal(size=500)), 300),
            ("first 32 NaN", lambda: np.concatenate([np.full(32, np.nan), rng.normal(size=100)]), 100),
            ("all NaN but 1", lambda: np.concatenate([np.full(99, np.nan), [3.0]]), 3),
            ("all equal", lambda: np.full(500, 2.5), 3),
            ("60% zeros", lambda: rng.permutation(np.concatenate([np.zeros(300), rng.normal(size=200)])), 200),
            ("integers", lambda: np.round(rng.normal(size=500) * 3), 300),
            ("two values", lambda: rng.permutation(np.concatenate([np.full(250, -1.0), np.full(250, 1.0)])), 50),
            ("+-inf", lambda: rng.permutation(np.concatenate([np.full(200, np.inf), np.full(200, -np.inf), rng.normal(size=100)])), 50),
            ("mixed zeros", lambda: rng.permutation(np.concatenate([np.zeros(100), -np.zeros(100), rng.normal(size=64) * 1e-30])), 50),
            ("n=64", lambda: rng.normal(size=64), 300), ("n=65", lambda: rng.normal(size=65), 300),
            ("sorted", lambda: np.sort(rng.normal(size=500)), 100),
            ("denormals", lambda: rng.normal(size=300) * 1e-42, 100)):
        stats = []
        for _ in range(cnt):
            _check(gen(), stats)
        print(f"{name:18s} passes mean {np.mean(stats):5.2f}  p99 {np.percentile(stats, 99):5.1f}  max {np.max(stats)}")


if __name__ == "__main__":
    main()
