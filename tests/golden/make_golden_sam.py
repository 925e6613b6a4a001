"""Generates tests/golden/golden_sam_v1.json from the REFERENCE'S OWN Mappinghandler (oracle/_ref/libhrm_ref_sam.so:
/root/reference/src/gpu/mappinghandler.cu compiled where it lies, see oracle/ref_shim_sam.cpp).  Build container only:

    make -C oracle ref && python tests/golden/make_golden_sam.py

Per case of tests/samcase.py: sha256 of the reference's SAM text, its size, the per-read values after the
recalculation (scores, conversion counts, flags) and a few literal record lines.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
import samcase  # noqa: E402


def main():
    port = po.Oracle("port")
    G = {"generator": "tests/golden/make_golden_sam.py", "source": "oracle/_ref/libhrm_ref_sam.so (reference sources)",
         "cases": {}}
    for name in samcase.CASES:
        case = samcase.make(port, name)
        sam, per = samcase.reference_sam(po, case)
        lines = sam.split(b"\n")
        n = len(case["lens"])
        rec = lines[n + 2:n + 2 + n]
        pick = sorted(set([0, 1, 2, n // 2, n - 1] + [i for i in range(n) if b"YZ:A:<->" in rec[i]][:3]))
        G["cases"][name] = {"sha256": hashlib.sha256(sam).hexdigest(), "bytes": len(sam), "per_read": per.tolist(),
                            "lines": {str(i): rec[i].decode() for i in pick}}
        print(name, len(sam), "bytes,", sum(b"YZ:A:<->" in x for x in rec), "minus-strand records,",
              int((case["mapped"]["orientation"] != 3).sum()), "mapped of", n)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_sam_v1.json")
    with open(out, "w") as f:
        json.dump(G, f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
