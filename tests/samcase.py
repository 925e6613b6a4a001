"""Deterministic inputs for the V4 / O1 (recalculation, alignment choice, MAPQ, SAM text) parity tests:
one pass on pre-converted input = what the reference's Mappinghandler is given.  The MappedReads
come from the oracle's seeding pass and are then perturbed so that every branch of printtoSAM is taken:
flipped orientations (alignment 1 wins -> YZ:A:<->), hits in the first windows of a chromosome (rc_ref beyond
the chromosome end), the short last window, neighbouring windows (clipped alignments), unmapped."""
import numpy as np

from hashreadmapper_b200 import synth

CASES = {
    # name: (chromosome lengths, n reads, read length, error rate, indel fraction, stage-V conversion, seed)
    "ct150": ([40007, 21013], 1200, 150, 0.02, 0.0, 1, 31),
    "ga150": ([30011, 9001], 900, 150, 0.02, 0.0, 2, 32),
    "ct250": ([50021], 700, 250, 0.03, 0.1, 1, 33),
    # the reference as shipped: unconverted reads and genome seed the hit, stage V converts C->T (conv 0 below)
    "none100": ([20011], 500, 100, 0.03, 0.0, 0, 34),
}
NAMES = ["chrA", "chrB", "chrC"]


def make(port, name, w=128, k=16):
    lengths, n, L, err, indel, conv, seed = CASES[name]
    genome, off = synth.make_genome(lengths, seed=seed)
    reads, lens, _ = synth.make_reads(genome, off, n, L, error_rate=err, indel_frac=indel, seed=seed + 100,
                                      conversion_rate=1.0 if conv else 0.0)
    rng = np.random.Generator(np.random.PCG64(seed + 200))
    lens[3] = 40
    lens[11] = max(16, L - 37)
    gconv = conv
    rconv = 1 if conv else 0
    g = port.convert_ascii(genome, gconv)
    r = np.frombuffer(port.convert_ascii(reads.tobytes(), rconv), dtype=np.uint8).reshape(reads.shape).copy()
    mapped, _ = port.map_pass_refdir(g, off, r, lens, k=k, w=w)
    stride = w - k + 1
    clen = np.diff(off)
    idx = np.nonzero(mapped["orientation"] != 3)[0]
    for i in idx:
        u = rng.random()
        c = int(mapped["chromosomeId"][i])
        nwin = (int(clen[c]) + stride - 1) // stride
        if u < 0.08:      # wrong strand: alignment 1 (the reverse-complement query) wins
            mapped["orientation"][i] = 3 - mapped["orientation"][i]
        elif u < 0.14:    # neighbouring window
            wid = int(mapped["position"][i]) // stride + (1 if rng.random() < 0.5 else -1)
            mapped["position"][i] = min(max(wid, 0), nwin - 1) * stride
        elif u < 0.17:    # first two windows: rc_ref runs past the chromosome
            mapped["position"][i] = int(rng.integers(0, 2)) * stride
        elif u < 0.20:    # the (short) last window
            mapped["position"][i] = (nwin - 1) * stride
        elif u < 0.22:
            mapped["orientation"][i] = 3
            mapped["chromosomeId"][i] = 0
            mapped["position"][i] = 0
    unm = mapped["orientation"] == 3
    mapped["chromosomeId"][unm] = 0
    mapped["position"][unm] = 0
    mapped["hammingDistance"][unm] = 0
    mapped["shift"][unm] = 0
    return {"genome": g, "off": off, "names": NAMES[:len(lengths)], "reads": r, "lens": lens, "mapped": mapped,
            "conv": conv if conv else 1, "w": w, "k": k,
            # what the device path is given: unconverted text + the conversions of its single pass
            "raw_genome": genome, "raw_reads": reads, "rconv": rconv, "gconv": gconv}


def reference_sam(po, case):
    """the reference's own Mappinghandler on the case (G->A cases through the complement mirror)"""
    g, r = case["genome"], case["reads"]
    if case["conv"] == 2:
        g = po.complement_ascii(g)
        r = np.frombuffer(po.complement_ascii(r.tobytes()), dtype=np.uint8).reshape(r.shape)
    sam, per = po.ref_mapping_sam(g, case["off"], case["names"], r, case["lens"], case["mapped"], w=case["w"])
    if case["conv"] == 2:
        sam = po.sam_complement_text_columns(sam)
    return sam, per


def port_sam(po, port, case, with_header=True):
    n = len(case["lens"])
    return po.port_sam_format(port, [case["genome"]], case["off"], case["names"], [case["reads"]], case["lens"],
                              case["mapped"], np.zeros(n, np.int32), [case["conv"]], w=case["w"],
                              with_header=with_header)
