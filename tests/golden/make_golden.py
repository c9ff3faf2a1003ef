"""Build the golden fixtures of tests/golden/ from the reference's own test data (run in the dev container only).

    python tests/golden/make_golden.py /root/reference

Sources (pyresample/test/test_files/, read-only): the 800x800 text grids used by test_kd_tree.py:488-609 and the
real SSMIS swath of test_swath.py.  They are data, not code; they are stored compactly (bit-packed masks,
compressed arrays) so that the oracle can be pinned on any machine without the reference tree.
"""
import os
import sys

import numpy as np


def main(ref_root):
    src = os.path.join(ref_root, "pyresample", "test", "test_files")
    out = os.path.dirname(os.path.abspath(__file__))
    grids = {}
    for name in ("mask_test_nearest_mask", "mask_test_mask", "mask_test_fill_value", "mask_test_full_fill"):
        arr = np.fromfile(os.path.join(src, name + ".dat"), sep=" ").reshape((800, 800))
        assert set(np.unique(arr)) <= {0.0, 1.0}
        grids[name] = np.packbits(arr.astype(bool))
        grids[name + "_sum"] = np.array(arr.sum())
    for name in ("mask_test_nearest_data", "mask_test_data"):
        arr = np.fromfile(os.path.join(src, name + ".dat"), sep=" ").reshape((800, 800))
        grids[name] = arr
    np.savez_compressed(os.path.join(out, "kd_tree_mask_grids.npz"), **grids)
    sw = np.load(os.path.join(src, "ssmis_swath.npz"))["data"]
    np.savez_compressed(os.path.join(out, "ssmis_swath.npz"), data=sw)
    for f in ("kd_tree_mask_grids.npz", "ssmis_swath.npz"):
        print(f, os.path.getsize(os.path.join(out, f)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
