"""Throughput of the device read ingestion (hrm_ingest_reads) next to the reference's own parser.

    python tools/bench_ingest.py [n_reads]

FASTQ text of n 150 bp reads resident in HBM -> normalised ASCII rows + lengths.  Algorithmic bytes: the text is read
twice (line index, copy) and the rows are written once; the kernels read the text a third time (newline count per
tile before the positions can be written).  CPU arm: forEachReadInFile (kseqpp) on the same file
through oracle/_ref, one host thread as in the reference."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hashreadmapper_b200.api as api  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    L = 150
    rng = np.random.Generator(np.random.PCG64(3))
    seq = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.choice(5, size=(n, L), p=[0.2495, 0.2495, 0.2495, 0.2495, 0.002])]
    rec = np.empty((n, 8 + 1 + L + 1 + 2 + L + 1), dtype=np.uint8)  # "@rNNNNNN\nSEQ\n+\nQUAL\n"
    ids = np.char.zfill(np.arange(n).astype("U7"), 7).astype("S7")
    rec[:, 0] = ord("@")
    rec[:, 1:8] = np.frombuffer(ids.tobytes(), dtype=np.uint8).reshape(n, 7)
    rec[:, 8] = 10
    rec[:, 9:9 + L] = seq
    rec[:, 9 + L] = 10
    rec[:, 10 + L] = ord("+")
    rec[:, 11 + L] = 10
    rec[:, 12 + L:12 + 2 * L] = ord("I")
    rec[:, 12 + 2 * L] = 10
    text = rec.reshape(-1)
    d_text = torch.from_numpy(text).cuda()
    pitch = 160
    for _ in range(3):
        rows, lens, amb, _ = api.ingest_reads(d_text, pitch, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 5
    for _ in range(steps):
        rows, lens, amb, _ = api.ingest_reads(d_text, pitch, n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert rows.shape[0] == n and int(lens.min()) == L and int(lens.max()) == L
    alg = 2.0 * text.size + n * pitch + 4.0 * n
    peak = 6457.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    line = {"metric": "reads ingested/sec", "value": n / (ms / 1e3), "unit": "reads/s", "ms_per_step": ms,
            "config": {"workload": "%d x 150bp FASTQ records (%.0f MB of text) resident in HBM" % (n, text.size / 1e6)},
            "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / 1e3) / 1e9 / peak,
                         "note": "algorithmic bytes = 2 x text (line index, copy) + rows + lengths; the kernels read the "
                                 "text a third time (newline count per tile)"},
            "ambiguous_reads": int(amb.sum().item())}
    try:
        from oracle.pyoracle import Oracle, have_ref
        if have_ref():
            ref = Oracle("ref")
            sub = min(n, 500_000)
            with tempfile.NamedTemporaryFile(suffix=".fastq", delete=False) as f:
                f.write(rec[:sub].tobytes())
                path = f.name
            t0 = time.perf_counter()
            r, l = ref.read_file(path, pitch, sub)
            dt = time.perf_counter() - t0
            os.unlink(path)
            line["cpu_baseline"] = {"value": sub / dt, "unit": "reads/s", "cores": 1, "kind": "reference",
                                    "sample": "%d reads parsed by forEachReadInFile (kseqpp) from a file in the page "
                                              "cache, without the normalisation pass" % sub}
    except Exception as e:
        line["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (e,)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
