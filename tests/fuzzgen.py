"""Seeded read generators for the parity fuzz tests (small, pure Python)."""
import random

ACGT = b"ACGT"


def rand_seq(rng, n, alphabet=ACGT):
    return bytes(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, rate, alphabet=b"ACGTNacgt"):
    if rate <= 0:
        return s
    b = bytearray(s)
    for i in range(len(b)):
        if rng.random() < rate:
            b[i] = rng.choice(alphabet)
    return bytes(b)


def planted_read(rng, length, dr_len=None, spacer_lo=26, spacer_hi=50, jitter=3, lead=None, sub_rate=0.0, dr=None):
    """A read of `length` bases cut from background + DR/spacer/DR/... array."""
    if dr is None:
        dr_len = dr_len or rng.randint(23, 47)
        dr = rand_seq(rng, dr_len)
    base_sp = rng.randint(spacer_lo, spacer_hi)
    arr = bytearray()
    lead = rng.randint(0, 80) if lead is None else lead
    arr += rand_seq(rng, lead)
    while len(arr) < length + 100:
        arr += mutate(rng, dr, sub_rate, ACGT)
        sp = max(1, base_sp + rng.randint(-jitter, jitter))
        arr += rand_seq(rng, sp)
    off = rng.randint(0, 60)
    return bytes(arr[off:off + length])


def microsat_read(rng, length):
    unit = rand_seq(rng, rng.randint(1, 6))
    s = (unit * (length // len(unit) + 2))[:length]
    return mutate(rng, s, rng.choice([0.0, 0.0, 0.01, 0.05]), ACGT)


def fuzz_read(rng, max_len=400):
    kind = rng.random()
    length = rng.choice([0, 1, 30, 57, 58, 59, 60, 75, 100, 101, 150, 150, 150, 250]) if rng.random() < 0.7 else rng.randint(0, max_len)
    if kind < 0.25:
        s = rand_seq(rng, length)
    elif kind < 0.80:
        s = planted_read(rng, length, sub_rate=rng.choice([0.0, 0.0, 0.01, 0.03])) if length else b""
    elif kind < 0.90:
        s = microsat_read(rng, length) if length else b""
    else:
        # tandem: repeat unit of DR-like length back to back, or short spacers
        s = planted_read(rng, length, spacer_lo=rng.randint(1, 30), spacer_hi=rng.randint(30, 60), jitter=rng.randint(0, 12)) if length else b""
    if rng.random() < 0.15:
        s = mutate(rng, s, rng.choice([0.005, 0.02]), b"NnacgtRY")
    return s


def dr_like_patterns(rng, n, lo=23, hi=47, both_strands=True):
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    out = []
    for _ in range(n):
        p = rand_seq(rng, rng.randint(lo, hi))
        out.append(p)
        if both_strands:
            out.append(p.translate(comp)[::-1])
    return out


def uss_case(rng):
    """Input of ReadHolder::updateStartStops as WorkHorse produces it: a read that carries 1-4 copies of a DR
    (found repeats = trimmed copies, as seed extension leaves them), often a partial copy at either end, and the
    group's consensus DR with the offset of the found repeat inside it.  -> (read, start/stops, front offset, DR)"""
    dr = rand_seq(rng, rng.randint(23, 47))
    trim_l, trim_r = rng.randint(0, 6), rng.randint(0, 6)
    parts, ss, pos = [], [], 0
    if rng.random() < 0.7:
        k = rng.randint(1, len(dr))
        piece = mutate(rng, dr[len(dr) - k:], 0.03, b"ACGTN")
    else:
        piece = rand_seq(rng, rng.randint(0, 40))
    parts.append(piece)
    pos += len(piece)
    for _ in range(rng.randint(1, 4)):
        sp = rand_seq(rng, rng.randint(26, 50))
        parts.append(sp)
        pos += len(sp)
        parts.append(mutate(rng, dr, 0.02, b"ACGTN"))
        ss += [pos + trim_l, pos + len(dr) - 1 - trim_r]
        pos += len(dr)
    sp = rand_seq(rng, rng.randint(0, 50))
    parts.append(sp)
    if rng.random() < 0.7:
        parts.append(mutate(rng, dr[: rng.randint(1, len(dr))], 0.03, b"ACGTN"))
    seq = b"".join(parts)
    if rng.random() < 0.1:
        seq = seq[: max(ss[-1] + 1, len(seq) - rng.randint(0, 30))]
    front = trim_l if rng.random() < 0.7 else rng.randint(-4, 12)
    if rng.random() < 0.1:
        dr = dr + rand_seq(rng, rng.randint(1, 10))
    if rng.random() < 0.05:
        seq = seq.lower() if rng.random() < 0.5 else seq.replace(b"A", b"R", 1)
    return seq, ss, front, dr


def consensus_case(rng, n_slaves=None, read_len=(100, 150), alphabet=b"ACGT"):
    """A DR group as WorkHorse::parseGroupedDRs hands it to the Aligner: DR 0 is the master (the longest), the others are
    variants of it -- trimmed or grown at either end, with substitutions, some on the other strand, some of them their own
    reverse complement (forward and reverse scores equal: the extendSlaveDR path), one unrelated (alignment fails) -- and
    every DR has a few reads that carry it one to three times (first repeat sometimes a partial one)."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    master = rand_seq(rng, rng.randint(34, 46))
    drs = [master]
    for _ in range(n_slaves if n_slaves is not None else rng.randint(2, 10)):
        kind = rng.random()
        if kind < 0.1:
            v = rand_seq(rng, rng.randint(23, 32))                                  # unrelated
        elif kind < 0.25:
            half = master[rng.randint(0, 6): rng.randint(14, 18)]
            v = half + half.translate(comp)[::-1]                                   # its own reverse complement
        else:
            v = master[rng.randint(0, 6): len(master) - rng.randint(0, 6)]
            v = mutate(rng, v, rng.choice([0, 0, 0.03, 0.08]), b"ACGT")
            if rng.random() < 0.3:
                v = rand_seq(rng, rng.randint(0, 2)) + v + rand_seq(rng, rng.randint(0, 2))
            if rng.random() < 0.3:
                v = v.translate(comp)[::-1]
        if len(v) >= 12 and len(v) < len(master) and v not in drs:
            drs.append(v)
    reads = []
    for d, dr in enumerate(drs):
        for _ in range(rng.randint(1, 4)):
            L = rng.randint(*read_len)
            seq = bytearray(rand_seq(rng, L, alphabet))
            ss = []
            pos = rng.randint(0, 12)
            if rng.random() < 0.3 and pos >= 0:                                       # a partial repeat in front (its tail only)
                cut = rng.randint(3, len(dr) - 2)
                part = dr[cut:]
                seq[0:len(part)] = part
                ss += [0, len(part) - 1]
                pos = len(part) + rng.randint(26, 36)
            for _k in range(rng.randint(1, 3)):
                if pos + len(dr) > L:
                    break
                seq[pos:pos + len(dr)] = dr
                ss += [pos, pos + len(dr) - 1]
                pos += len(dr) + rng.randint(26, 40)
            if not any(ss[k + 1] - ss[k] == len(dr) - 1 for k in range(0, len(ss), 2)):
                continue
            reads.append((bytes(seq[:L]), ss, d))
    if not any(r[2] == 0 for r in reads):                                           # the master always has a read
        L = read_len[1]
        seq = bytearray(rand_seq(rng, L))
        seq[40:40 + len(master)] = master
        reads.append((bytes(seq), [40, 40 + len(master) - 1], 0))
    return dict(reads=reads, drs=drs, array_len=4 * read_len[1])
