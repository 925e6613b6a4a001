"""Worker of test_gpu_mapper.py::test_fused_collection_paths (TEST-ONLY): maps one fixed data set and saves the
mapped reads, so that the test can compare runs whose kernel geometry was changed through the HRM_COLLECT_*
environment hooks (they are read once per process)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hashreadmapper_b200.api as api  # noqa: E402
from hashreadmapper_b200 import synth  # noqa: E402

out_path, min_hits = sys.argv[1], int(sys.argv[2])
# low-complexity genome: long T-rich stretches give buckets with hundreds of windows
rng = np.random.Generator(np.random.PCG64(5))
g = synth.ALPHABET[rng.choice(4, size=1_200_000, p=[0.15, 0.05, 0.1, 0.7]).astype(np.uint8)]
for _ in range(300):  # planted repeats: the same 400-mer at many places (many windows pass the filter)
    p = int(rng.integers(0, len(g) - 400))
    g[p:p + 400] = g[1000:1400]
genome, off = g.tobytes(), np.array([0, 700_000, len(g)], dtype=np.int64)
reads, lens, _ = synth.make_reads(genome, off, 6000, 150, error_rate=0.01, seed=9)
cfg = api.directional_config()
cfg.min_table_hits = min_hits
mp = api.Mapper(cfg)
mp.setGenome(genome, off)
out, st = mp.mapBatch(torch.from_numpy(reads).cuda(), torch.from_numpy(lens).cuda())
np.save(out_path, out.cpu().numpy())
print("values", st.num_values, "candidates", st.num_candidates, "mapped", st.num_mapped)
