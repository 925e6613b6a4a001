import random

import numpy as np


def rs(rng, n, alpha="ACGT"):
    return "".join(rng.choice(alpha) for _ in range(n)).encode()


def mutate(rng, s, sub, indel):
    out = bytearray()
    for ch in s:
        r = rng.random()
        if r < sub:
            out.append(rng.choice(b"ACGT"))
        elif r < sub + indel / 2:
            pass
        elif r < sub + indel:
            out.append(ch)
            out.append(rng.choice(b"ACGT"))
        else:
            out.append(ch)
    return bytes(out)


def rows(seqs, pitch=None, fill=0):
    """list of bytes -> (uint8 [n, pitch] with pitch % 16 == 0, int32 lengths)"""
    mx = max([len(s) for s in seqs] + [1])
    pitch = pitch or ((mx + 15) // 16) * 16
    a = np.full((len(seqs), pitch), fill, dtype=np.uint8)
    for i, s in enumerate(seqs):
        a[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    return a, np.array([len(s) for s in seqs], dtype=np.int32)


def pack_rows(port, seqs, pitch_words=None):
    mx = max([len(s) for s in seqs] + [1])
    pw = pitch_words or (mx + 15) // 16
    out = np.zeros((len(seqs), pw), dtype=np.uint32)
    for i, s in enumerate(seqs):
        e = port.encode_2bit(s)
        out[i, :len(e)] = e
    return out


def ssw_cases(seed, n):
    """(query, ref, maskLen) triples exercising byte/word mode, indels, 3-letter alphabets, N"""
    rng = random.Random(seed)
    out = []
    for it in range(n):
        mode = it % 8
        alpha = rng.choice(["ACGT", "AGT", "AT", "ACGTN", "AAAT"])
        if mode < 5:
            G = rs(rng, 600, alpha)
            p = rng.randint(100, 350)
            w = rng.choice([64, 128, 128, 256])
            ref = G[p:p + w]
            L = rng.choice([150, 150, 250, 100, 36])
            off = rng.randint(-L // 2, w // 2)
            q = mutate(rng, G[p + off:p + off + L], rng.choice([0, 0.01, 0.03, 0.1]),
                       rng.choice([0, 0, 0.003, 0.01, 0.05])) or b"A"
        elif mode == 5:
            q = rs(rng, rng.randint(1, 260), alpha)
            ref = rs(rng, rng.randint(1, 260), alpha)
        elif mode == 6:
            ref = rs(rng, rng.randint(120, 256), alpha)
            st = rng.randint(0, 20)
            q = ref[st:st + rng.randint(100, 250)]
            if rng.random() < 0.5:
                q = mutate(rng, q, 0.01, 0.005) or b"A"
        else:
            ref = rs(rng, rng.randint(10, 200), "A")
            q = rs(rng, rng.randint(10, 200), "AAAAC")
        ml = max(15, len(q) // 2) if rng.random() < 0.8 else rng.choice([15, 16, 30])
        out.append((q, ref, ml))
    return out
