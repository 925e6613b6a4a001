"""Synthetic genomes and bisulfite reads.  numpy only; used by tests and bench.

Reference genome: i.i.d. uniform over ACGT, seed 20240601.  Reads: seed 20240602, start uniform over
valid positions, strand + with p = 0.5; directional bisulfite model: a + read is C->T of the top
strand segment, a - read is C->T of the reverse complement of the segment; every C converts
(conversion rate 1.0 by default); substitution errors at `error_rate`, of which `indel_frac` are
1-bp indels; non-directional libraries add the two PCR-complement strands (G->A reads).
"""
import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def make_genome(lengths, seed=20240601):
    """-> (ascii bytes of all chromosomes concatenated, offsets int64 [n+1])"""
    rng = np.random.Generator(np.random.PCG64(seed))
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lengths)
    g = ALPHABET[rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)]
    return g.tobytes(), off


def revcomp(a):
    return _COMP[a[::-1]]


def convert(a, mode):
    a = a.copy()
    if mode == 1:
        a[a == ord("C")] = ord("T")
    elif mode == 2:
        a[a == ord("G")] = ord("A")
    return a


def make_reads(genome, off, n, length=150, error_rate=0.0, indel_frac=0.0, conversion_rate=1.0,
               nondirectional=False, seed=20240602, pitch=None):
    """-> (reads uint8 [n, pitch], lengths int32 [n], truth dict(chrom, pos, strand))"""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.frombuffer(genome, dtype=np.uint8)
    pitch = pitch or ((length + 15) // 16) * 16
    nchrom = len(off) - 1
    clen = np.diff(off)
    usable = np.maximum(clen - length - 2, 1)
    chrom = rng.choice(nchrom, size=n, p=usable / usable.sum())
    pos = (rng.random(n) * usable[chrom]).astype(np.int64)
    strand = rng.random(n) < 0.5  # True = +
    kind = np.zeros(n, dtype=np.int8)  # 0: C->T read (OT/OB), 1: G->A read (CTOT/CTOB)
    if nondirectional:
        kind = (rng.random(n) < 0.5).astype(np.int8)
    reads = np.full((n, pitch), 0, dtype=np.uint8)
    lens = np.full(n, length, dtype=np.int32)
    # vectorised gather of segments (+2 spare bases for deletions)
    idx = (off[chrom] + pos)[:, None] + np.arange(length + 2)[None, :]
    seg = g[np.minimum(idx, off[-1] - 1)]
    for i in range(n) if (indel_frac > 0 and error_rate > 0) else ():
        pass
    body = seg[:, :length].copy()
    # reverse strand: reverse complement of the segment
    rc = _COMP[body[:, ::-1]]
    body = np.where(strand[:, None], body, rc)
    # bisulfite conversion
    conv_mask = rng.random(body.shape) < conversion_rate
    ct = (body == ord("C")) & conv_mask & (kind[:, None] == 0)
    ga = (body == ord("G")) & conv_mask & (kind[:, None] == 1)
    body[ct] = ord("T")
    body[ga] = ord("A")
    # sequencing errors: substitutions
    if error_rate > 0:
        err = rng.random(body.shape) < error_rate * (1.0 - indel_frac)
        sub = ALPHABET[rng.integers(0, 4, size=body.shape, dtype=np.uint8)]
        body = np.where(err, sub, body)
        if indel_frac > 0:
            nind = rng.binomial(length, error_rate * indel_frac, size=n)
            for i in np.nonzero(nind)[0]:
                row = list(body[i])
                for _ in range(nind[i]):
                    p = int(rng.integers(1, len(row) - 1))
                    if rng.random() < 0.5:
                        del row[p]
                    else:
                        row.insert(p, int(ALPHABET[rng.integers(0, 4)]))
                row = (row + [ord("A")] * length)[:length]
                body[i] = np.array(row, dtype=np.uint8)
    reads[:, :length] = body
    return reads, lens, {"chrom": chrom.astype(np.int32), "pos": pos, "strand": strand, "kind": kind}


def human_like_lengths(total=3_100_000_000, n=24):
    """24 chromosome lengths with GRCh38-like proportions summing to `total` (each < 2^31)"""
    rel = np.array([248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47,
                    51, 156, 57], dtype=np.float64)[:n]
    ln = np.floor(rel / rel.sum() * total).astype(np.int64)
    ln[0] += total - ln.sum()
    return ln
