"""Generate tests/golden/*.npz: stage-by-stage inputs/outputs of the REFERENCE (oracle/_ref, scalar C
flavour, -O2 -fno-fast-math -ffp-contract=off) for a few blocks of each block type.  Needs
/root/reference (this container); the vectors are committed so that the oracle port and the CUDA
path stay pinned where the reference is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import refdata  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = [("q4_44k_2ch", 1.2, 4), ("q2_44k_2ch", 1.2, 4), ("abr128_44k_2ch", 0.8, 2), ("q8_48k_6ch", 0.6, 2)]


def main():
    os.makedirs(OUT, exist_ok=True)
    for config, seconds, keep in CASES:
        cap = refdata.capture(config, seconds)
        picked, seen = [], set()
        for seq in sorted(cap["blocks"]):
            b = cap["blocks"][seq]
            if "TAP_BLOCK" not in b:
                continue
            blk = b["TAP_BLOCK"][0]["data"]
            mode = int(blk[3]) | (int(blk[0]) << 1)
            if mode not in seen and seq > 2:
                seen.add(mode)
                picked.append(seq)
        picked = picked[-keep:]
        arrays = {"seqs": np.array(picked, np.int32)}
        for seq in picked:
            for tag, recs in cap["blocks"][seq].items():
                if tag in ("TAP_PACKET", "TAP_BLOCK_SYMS"):
                    continue
                for k, r in enumerate(recs):
                    key = "%d/%s/%d" % (seq, tag, k)
                    arrays[key] = r["data"]
                    arrays[key + "/a"] = np.array([r["a0"], r["a1"], r["a2"], r["a3"]], np.int32)
        path = os.path.join(OUT, config + ".npz")
        np.savez_compressed(path, **arrays)
        print(config, "blocks", picked, "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
