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
