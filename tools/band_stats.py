"""Band statistics of the SW trace-back stage on the bench workload (CPU, through the host harness).
Sizes the finish kernels: for each read, alignment 0 (true strand) and 1 (other strand):
refLen, readLen, final band width, band iterations."""
import ctypes as C
import collections
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hashreadmapper_b200 import synth

HH = os.path.join(ROOT, "tests", "host_harness")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(HH, "harness.cpp"),
                       "-o", os.path.join(HH, "libhh.so"), "-I", os.path.join(ROOT, "include")])
hh = C.CDLL(os.path.join(HH, "libhh.so"))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
err = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
indel = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
g, off = synth.make_genome([2_000_000])
reads, lens, truth = synth.make_reads(g, off, n, length=L, error_rate=err, indel_frac=indel)
G = np.frombuffer(g, dtype=np.uint8)
w, k = 128, 16
s = w - k + 1
stats = [collections.Counter(), collections.Counter()]
cells = [0, 0]
out = (C.c_int * 8)()
for i in range(n):
    pos = int(truth["pos"][i])
    strand = bool(truth["strand"][i])
    # best window ~ the one covering most of the read
    wid = max(0, (pos + L // 2 - w // 2 + s // 2) // s)
    win = G[wid * s: wid * s + w]
    rd = reads[i, :L]
    if strand:
        winc = synth.convert(win, 1)
        q0 = synth.convert(rd, 1)
        q1 = synth.convert(synth.revcomp(rd), 1)
    else:
        winc = synth.convert(win, 2)
        q0 = synth.convert(synth.revcomp(rd), 2)
        q1 = synth.convert(rd, 2)
    for a, q in enumerate((q0, q1)):
        hh.hh_sw_band_stats(q.tobytes(), L, winc.tobytes(), len(winc), max(15, L // 2), out)
        score, refLen, readLen, band, iters = out[0], out[1], out[2], out[3], out[4]
        stats[a][(band, iters)] += 1
        bb = abs(refLen - readLen) + 1
        for it in range(iters):
            cells[a] += readLen * min(2 * bb + 1, refLen)
            bb *= 2
        if i < 10:
            print(a, "score", score, "refLen", refLen, "readLen", readLen, "band", band, "iters", iters)
for a in range(2):
    print("alignment", a, "avg banded cells", cells[a] / n)
    for key, c in sorted(stats[a].items()):
        print("   band %4d iters %d : %6d (%.1f%%)" % (key[0], key[1], c, 100.0 * c / n))
