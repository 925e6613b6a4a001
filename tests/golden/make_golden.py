"""Generates tests/golden/golden_v1.json from the REFERENCE'S OWN code (oracle/_ref/libhrm_ref.so, built
from /root/reference by oracle/Makefile).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

The fixture pins the oracle (tests/test_oracle_pin.py) and, through it, the CUDA path on boxes where
/root/reference does not exist.  Inputs are seeded; every value below is an output of reference code.
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Oracle  # noqa: E402
from util import rs, mutate, ssw_cases  # noqa: E402


def main():
    R = Oracle("ref")
    rng = random.Random(20240601)
    G = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libhrm_ref.so (reference sources)"}
    # P1 / P1d
    enc = []
    for _ in range(60):
        s = rs(rng, rng.randint(1, 200), "ACGTN")
        e = R.encode_2bit(s)
        enc.append({"seq": s.decode(), "words": [int(x) for x in e],
                    "rc_words": [int(x) for x in R.revcomp_2bit(e, len(s))],
                    "rc_ascii": R.revcomp_ascii(s).hex()})
    G["encode"] = enc
    # H1/H2
    mh = []
    for _ in range(60):
        s = rs(rng, rng.choice([150, 128, 250, 40, 16, 15]), rng.choice(["ACGT", "AGT", "ACT"]))
        k = rng.choice([16, 16, 12, 20, 32])
        if len(s) < 32 and k == 32:
            k = 16
        e = R.encode_2bit(s)
        sig, val = R.minhash_batch(e[None, :], np.array([len(s)]), k, 16)
        mh.append({"seq": s.decode(), "k": k, "sig": [int(x) for x in sig[0]], "valid": [int(x) for x in val[0]]})
    G["minhash"] = mh
    G["murmur64"] = [[x, R.murmur64(x)] for x in (0, 1, 2, 0xdeadbeef, 2**63, 2**64 - 1, 123456789012345)]
    # H3 / H3q
    n, H = 400, 4
    sig = np.random.RandomState(1).randint(0, 60, size=(n, H)).astype(np.uint64)
    val = (np.random.RandomState(2).rand(n, H) > 0.05).astype(np.uint8)
    q = np.random.RandomState(3).randint(0, 70, size=(50, H)).astype(np.uint64)
    qv = np.ones((50, H), np.uint8)
    qv[::7] = 0
    tabs = {}
    for cap in (65535, 3):
        T = R.tables_build(sig, val, None, cap)
        num, off, vals = R.tables_query(T, q, qv)
        tabs[str(cap)] = {"num": num.tolist(), "values": vals.tolist()}
        R.tables_free(T)
    G["tables"] = {"sig_seed": 1, "valid_seed": 2, "query_seed": 3, "n": n, "H": H, "results": tabs}
    # S2
    wl = []
    for _ in range(200):
        a = (rng.randint(0, 500), rng.randint(500, 2000), rng.randint(500, 1900), rng.choice([64, 128, 256]),
             rng.randint(0, 130))
        wl.append([list(a), list(R.window_location(*a))])
    G["window_location"] = wl
    # S3
    shd = []
    for it in range(150):
        Lc = rng.choice([150, 100, 250, 36])
        La = Lc + rng.randint(-1, 140)
        A = rs(rng, La, "AGT" if it % 2 else "ACGT")
        if it % 3 and La >= Lc:
            st = rng.randint(0, La - Lc)
            c = bytearray(A[st:st + Lc])
            for _ in range(rng.randint(0, 10)):
                c[rng.randrange(Lc)] = rng.choice(b"ACGT")
            c = bytes(c)
            if it % 6 == 1:
                c = R.revcomp_ascii(c)
        else:
            c = rs(rng, Lc)
        res = R.shd(R.encode_2bit(A), La, R.encode_2bit(c), Lc, 0.05)
        shd.append({"anchor": A.decode(), "cand": c.decode(), "rate": 0.05, "result": list(res)})
    G["shd"] = shd
    # V2
    sw = []
    for q_, r_, ml in ssw_cases(20240602, 400):
        al, cig = R.ssw_align(q_, r_, ml)
        if al[0] == 0:
            continue  # degenerate: undefined in the reference
        sw.append({"q": q_.decode(), "r": r_.decode(), "mask": ml, "al": list(al), "cigar": cig})
    G["ssw"] = sw
    # V3
    ed = []
    for it in range(150):
        a = rs(rng, rng.randint(1, 200), "AGT")
        b = mutate(rng, a, 0.05, 0.05) if it % 2 else rs(rng, rng.randint(1, 200), "AGT")
        b = b or b"G"
        ed.append([a.decode(), b.decode(), R.edit_distance_nw(a, b)])
    G["edit"] = ed
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")
    with open(out, "w") as f:
        json.dump(G, f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
