"""GPU experiment: the fused collection on the human-size index under different kernel geometries (env hooks are read
per call).  Prints per-setting filter-stage ms per 1 M reads, ids enumerated / skipped, reads handed to the block kernel."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hashreadmapper_b200.api as api  # noqa: E402
from hashreadmapper_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
gbp = int(sys.argv[2]) if len(sys.argv) > 2 else 3_100_000_000
lengths = [int(x) for x in synth.human_like_lengths(gbp, 24)]
genome, off = synth.make_genome(lengths, seed=20240601)
reads, lens, _ = synth.make_reads(genome, off, n, 150, error_rate=0.01, seed=20240603)
mp = api.Mapper(api.directional_config())
t0 = time.time()
mp.setGenome(genome, off)
print("index built in %.1f s" % (time.time() - t0), flush=True)
d_reads, d_lens = torch.from_numpy(reads).cuda(), torch.from_numpy(lens).cuda()
settings = [{}] + [dict(e.split("=") for e in s.split(",")) for s in sys.argv[3:]]
ref = None
for env in settings:
    for k in list(os.environ):
        if k.startswith("HRM_COLLECT") and k != "HRM_COLLECT_DEBUG":
            del os.environ[k]
    os.environ.update(env)
    mp.mapBatch(d_reads, d_lens, want_stats=False)
    i0 = mp.info()
    mp.setProfiling(True)
    mp.stageTimes()
    out, st = mp.mapBatch(d_reads, d_lens, want_stats=True)
    torch.cuda.synchronize()
    stg = mp.stageTimes()
    mp.setProfiling(False)
    i1 = mp.info()
    o = out.cpu().numpy()
    if ref is None:
        ref = o
    print(env or "default", "filter %.1f ms" % stg["filter"][0], "enum %.0f skipped %.0f per read-pass, block-kernel reads %d (%.3f %%), same=%s"
          % ((i1.collect_ids_counted - i0.collect_ids_counted) / (2.0 * n), (i1.collect_ids_skipped - i0.collect_ids_skipped) / (2.0 * n),
             i1.collect_reads_block_kernel - i0.collect_reads_block_kernel,
             100.0 * (i1.collect_reads_block_kernel - i0.collect_reads_block_kernel) / (2.0 * n), bool((o == ref).all())), flush=True)
