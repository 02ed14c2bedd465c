"""Seeded synthetic loci of the shapes BASELINE.json names: UCE-like gamma
rates, 30 % invariant (exactly zero) sites, optional NaN (culled) sites."""
import numpy as np

EPOCHS = [[32, 42], [88, 98], [95, 105], [164, 174]]   # README.rst:135-136
TIMES = [37, 93, 100, 170]


def uce_rates(rng, n, T, zero_frac=0.3, nan_frac=0.0):
    lam = rng.gamma(0.5, 4.0 / T, size=n)          # mean rate 2/T: PI peaks 1/(4 lam) inside the grid
    lam[rng.random(n) < zero_frac] = 0.0
    if nan_frac:
        lam[rng.random(n) < nan_frac] = np.nan
    return lam


BLOCK_LOCI = 8192          # large uniform batches are generated in blocks of loci, one seeded stream per block
BLOCKED_ABOVE = 50_000_000  # sites


def csr_uniform(seed, n_loci, n_sites, T, nan_frac=0.0, out=None, loci=None, threads=None):
    """Uniform loci.  Small batches come from one stream (seed); batches above BLOCKED_ABOVE sites are defined block by
    block (stream (seed, block index) for loci [8192 b, 8192 (b+1))), so that they can be generated on several threads
    -- numpy releases the GIL inside Generator -- and any range of loci `loci=(l0, l1)` can be built on its own:
    the rates of a locus do not depend on who generates it.  Returns (offsets, rates) of the requested range."""
    l0, l1 = (0, n_loci) if loci is None else loci
    if n_loci * n_sites <= BLOCKED_ABOVE:
        rng = np.random.default_rng(seed)
        rates = uce_rates(rng, n_loci * n_sites, T, nan_frac=nan_frac)[l0 * n_sites:l1 * n_sites]
        return np.arange(l1 - l0 + 1, dtype=np.int64) * n_sites, rates
    import os
    from concurrent.futures import ThreadPoolExecutor
    rates = np.empty((l1 - l0) * n_sites, dtype=np.float64) if out is None else out
    blocks = range(l0 // BLOCK_LOCI, (l1 + BLOCK_LOCI - 1) // BLOCK_LOCI)

    def fill(b):
        b0, b1 = b * BLOCK_LOCI, min(n_loci, (b + 1) * BLOCK_LOCI)
        lam = uce_rates(np.random.default_rng([seed, b]), (b1 - b0) * n_sites, T, nan_frac=nan_frac)
        lo, hi = max(b0, l0), min(b1, l1)
        rates[(lo - l0) * n_sites:(hi - l0) * n_sites] = lam[(lo - b0) * n_sites:(hi - b0) * n_sites]

    threads = threads or max(1, min(16, (os.cpu_count() or 2) - 1))
    with ThreadPoolExecutor(max_workers=threads) as pool:
        list(pool.map(fill, blocks))
    return np.arange(l1 - l0 + 1, dtype=np.int64) * n_sites, rates


def csr_ragged(seed, n_loci, smin, smax, mean_sites, T, nan_frac=0.0):
    """log-uniform site counts in [smin, smax] rescaled to the requested mean (config 3)."""
    rng = np.random.default_rng(seed)
    s = np.exp(rng.uniform(np.log(smin), np.log(smax), size=n_loci))
    s = np.clip(np.round(s * (mean_sites / s.mean())), smin, smax).astype(np.int64)
    offsets = np.zeros(n_loci + 1, dtype=np.int64)
    np.cumsum(s, out=offsets[1:])
    rates = uce_rates(rng, int(offsets[-1]), T, nan_frac=nan_frac)
    return offsets, rates


CONFIGS = {
    # name: (builder kwargs, T)  -- BASELINE.json configs[1..4]
    "cfg2": dict(kind="uniform", seed=2, n_loci=5000, n_sites=500, T=1000),
    "cfg3": dict(kind="ragged", seed=3, n_loci=50000, smin=200, smax=5000, mean_sites=1000, T=4096),
    "cfg4": dict(kind="uniform", seed=4, n_loci=1000, n_sites=1000000, T=10000),
    "cfg5": dict(kind="uniform", seed=5, n_loci=500000, n_sites=800, T=2048),
}


def build(name, scale=1.0, loci=None, threads=None):
    """-> (offsets, rates, T).  `scale` < 1 shrinks the locus count (bounded CPU samples); `loci=(l0, l1)` builds only
    that range of a uniform batch."""
    c = dict(CONFIGS[name])
    kind, T = c.pop("kind"), c.pop("T")
    c["n_loci"] = max(1, int(round(c["n_loci"] * scale)))
    if kind == "uniform":
        off, r = csr_uniform(T=T, loci=loci, threads=threads, **c)
    else:
        off, r = csr_ragged(T=T, **c)
    return off, r, T
