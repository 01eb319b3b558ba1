"""Seeded synthetic read sets with planted CRISPR arrays (BASELINE.json configs 2-5).

A "genome" is uniform i.i.d. A/C/G/T background with CRISPR arrays spliced in at random places:
``n_dr_types`` direct repeats of length U[23,47] (rejected when low-complexity or when a 3-mer exceeds
23% of the repeat), each forming one array of 20-200 spacers of length U[26,50] with a per-array
jitter of at most 3.  Reads are sampled uniformly from the genome, from both strands, with 0.1%
substitutions and 0.05% N.  Everything is a pure function of the seed (numpy PCG64).
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def _rand_bases(rng, n):
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def _dr_ok(dr):
    n = len(dr)
    counts = np.bincount(dr, minlength=256)
    if counts.max() > int(n * 0.75):
        return False
    codes = (dr >> 1) & 3
    k3 = codes[:-2].astype(np.int64) * 16 + codes[1:-1] * 4 + codes[2:]
    return np.bincount(k3, minlength=64).max() / float(len(k3)) <= 0.23


def make_dr_types(rng, n_types=50, lo=23, hi=47):
    out = []
    while len(out) < n_types:
        dr = _rand_bases(rng, int(rng.integers(lo, hi + 1)))
        if _dr_ok(dr):
            out.append(dr)
    return out


def make_genome(seed, n_dr_types=50, array_fraction=0.01, min_spacers=20, max_spacers=200, sp_lo=26, sp_hi=50):
    """Returns (genome uint8[], list of DR byte strings, array_base_count)."""
    rng = np.random.default_rng(seed)
    drs = make_dr_types(rng, n_dr_types)
    arrays = []
    for dr in drs:
        n_sp = int(rng.integers(min_spacers, max_spacers + 1))
        base = int(rng.integers(sp_lo + 3, sp_hi - 2))
        parts = []
        for _ in range(n_sp):
            parts.append(dr)
            parts.append(_rand_bases(rng, base + int(rng.integers(-3, 4))))
        parts.append(dr)
        arrays.append(np.concatenate(parts))
    n_array = sum(len(a) for a in arrays)
    n_bg = int(n_array * (1.0 - array_fraction) / array_fraction)
    cuts = np.sort(rng.integers(0, n_bg, size=len(arrays)))
    bg = _rand_bases(rng, n_bg)
    parts, prev = [], 0
    for c, a in zip(cuts, arrays):
        parts.append(bg[prev:c])
        parts.append(a)
        prev = c
    parts.append(bg[prev:])
    return np.concatenate(parts), [d.tobytes() for d in drs], n_array


def _mutate_inplace(rng, flat, sub_rate, n_rate):
    n = flat.size
    if sub_rate > 0:
        k = rng.binomial(n, sub_rate)
        pos = rng.integers(0, n, size=k)
        flat[pos] = _ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]
    if n_rate > 0:
        k = rng.binomial(n, n_rate)
        flat[rng.integers(0, n, size=k)] = ord("N")


def sample_fixed(genome, n_reads, read_len, seed, sub_rate=0.001, n_rate=0.0005, chunk=1 << 20, out=None):
    """n_reads x read_len reads -> (bases uint8[n_reads*read_len], offsets uint64[n_reads+1])."""
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty(n_reads * read_len, dtype=np.uint8)
    ar = np.arange(read_len, dtype=np.int64)
    rar = ar[::-1].copy()
    hi = len(genome) - read_len
    for lo_i in range(0, n_reads, chunk):
        m = min(chunk, n_reads - lo_i)
        starts = rng.integers(0, hi, size=m, dtype=np.int64)
        rev = rng.integers(0, 2, size=m, dtype=np.uint8).astype(bool)
        idx = starts[:, None] + np.where(rev[:, None], rar[None, :], ar[None, :])
        blk = genome[idx]
        blk[rev] = _COMP[blk[rev]]
        flat = blk.reshape(-1)
        _mutate_inplace(rng, flat, sub_rate, n_rate)
        out[lo_i * read_len:(lo_i + m) * read_len] = flat
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return out, offsets


def sample_variable(genome, n_reads, len_lo, len_hi, seed, sub_rate=0.001, n_rate=0.0005, chunk=1 << 14):
    """Long reads with length U[len_lo, len_hi] -> (bases, offsets)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(len_lo, len_hi + 1, size=n_reads, dtype=np.int64)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens).astype(np.uint64)
    out = np.empty(int(offsets[-1]), dtype=np.uint8)
    for lo_i in range(0, n_reads, chunk):
        m = min(chunk, n_reads - lo_i)
        l = lens[lo_i:lo_i + m]
        starts = rng.integers(0, len(genome) - len_hi, size=m, dtype=np.int64)
        rev = rng.integers(0, 2, size=m, dtype=np.uint8).astype(bool)
        tot = int(l.sum())
        first = np.zeros(m, dtype=np.int64)
        first[1:] = np.cumsum(l)[:-1]
        rid = np.repeat(np.arange(m), l)
        within = np.arange(tot, dtype=np.int64) - first[rid]
        pos = np.where(rev[rid], starts[rid] + l[rid] - 1 - within, starts[rid] + within)
        blk = genome[pos]
        rmask = rev[rid]
        blk[rmask] = _COMP[blk[rmask]]
        _mutate_inplace(rng, blk, sub_rate, n_rate)
        b0 = int(offsets[lo_i])
        out[b0:b0 + tot] = blk
    return out, offsets


def config2(n_reads=10_000_000, read_len=150, seed=20242):
    """BASELINE.json configs[1]: 10M x 150 bp, 50 planted DR types, 1% array bases."""
    genome, drs, _ = make_genome(seed)
    bases, offsets = sample_fixed(genome, n_reads, read_len, seed + 1000)
    return bases, offsets, drs


def config3(n_reads=2_000_000, seed=20243, len_lo=1000, len_hi=10000):
    """BASELINE.json configs[2]: long reads 1-10 kb."""
    genome, drs, _ = make_genome(seed, array_fraction=0.03, min_spacers=60, max_spacers=200)
    bases, offsets = sample_variable(genome, n_reads, len_lo, len_hi, seed + 1000)
    return bases, offsets, drs


def pattern_set(n_patterns, seed=20245, lo=23, hi=47):
    """BASELINE.json configs[4]: P/2 random DR-like strings + their reverse complements."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_patterns // 2):
        p = _rand_bases(rng, int(rng.integers(lo, hi + 1)))
        out.append(p.tobytes())
        out.append(_COMP[p[::-1]].tobytes())
    return out


def plant_patterns(bases, offsets, patterns, fraction, seed):
    """Overwrites a random window of `fraction` of the (fixed-length) reads with one pattern each."""
    rng = np.random.default_rng(seed)
    n = len(offsets) - 1
    read_len = int(offsets[1] - offsets[0])
    pick = np.flatnonzero(rng.random(n) < fraction)
    for r in pick:
        p = patterns[int(rng.integers(0, len(patterns)))]
        at = int(rng.integers(0, read_len - len(p) + 1))
        b0 = int(offsets[r]) + at
        bases[b0:b0 + len(p)] = np.frombuffer(p, dtype=np.uint8)
    return pick


def sample_fixed_torch(genome, n_reads, read_len, seed, device, sub_rate=0.001, n_rate=0.0005, chunk=1 << 21):
    """Same recipe as sample_fixed but gathered on `device` with torch (fast enough for 10M+ reads inside a
    benchmark run).  Deterministic for a given seed / torch build; not byte-identical to sample_fixed.
    Returns (bases uint8 tensor on device, offsets int64 tensor on device)."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    dev = torch.device(device)
    G = torch.from_numpy(genome).to(dev)
    comp = torch.from_numpy(_COMP).to(dev)
    acgt = torch.from_numpy(_ACGT.copy()).to(dev)
    out = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, dtype=torch.int64, device=dev)
    hi = len(genome) - read_len
    for lo_i in range(0, n_reads, chunk):
        m = min(chunk, n_reads - lo_i)
        starts = torch.randint(0, hi, (m,), generator=g, dtype=torch.int64).to(dev)
        rev = torch.randint(0, 2, (m,), generator=g, dtype=torch.int64).to(dev).bool()
        idx = torch.where(rev[:, None], starts[:, None] + (read_len - 1) - ar[None, :], starts[:, None] + ar[None, :])
        blk = G[idx]
        blk = torch.where(rev[:, None], comp[blk.long()], blk)
        flat = blk.reshape(-1)
        n = flat.numel()
        k = int(torch.binomial(torch.tensor(float(n)), torch.tensor(sub_rate), generator=g).item()) if sub_rate > 0 else 0
        if k:
            pos = torch.randint(0, n, (k,), generator=g, dtype=torch.int64).to(dev)
            flat[pos] = acgt[torch.randint(0, 4, (k,), generator=g, dtype=torch.int64).to(dev)]
        k = int(torch.binomial(torch.tensor(float(n)), torch.tensor(n_rate), generator=g).item()) if n_rate > 0 else 0
        if k:
            flat[torch.randint(0, n, (k,), generator=g, dtype=torch.int64).to(dev)] = ord("N")
        out[lo_i * read_len:(lo_i + m) * read_len] = flat
    offsets = torch.arange(n_reads + 1, dtype=torch.int64, device=dev) * read_len
    return out, offsets
