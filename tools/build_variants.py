"""Kernel experiments: build variants of the library (macro settings) next to the product's, units float_8_3 and
float_16_0 only (the C3 / C4 search kernels)."""
import os
import subprocess
import sys

sys.path.insert(0, ".")
VARIANTS = {
    "v1": "-DPRS_SEED=0",
    "v2": "-DPRS_SEED=0 -DPRS_ROW_BATCH=1",
    "v3": "-DPRS_SEED=1 -DPRS_COOP=1",
    "v4": "-DPRS_SEED=1 -DPRS_ROW_BATCH=8",
}
for name, flags in VARIANTS.items():
    env = dict(os.environ, PRS_B200_LIBRARY=os.path.abspath("pyresample_b200/libprs_%s.so" % name), PRS_NVCC_FLAGS=flags,
               PRS_BUILD_ONLY="float_8_3,float_16_0")
    subprocess.check_call([sys.executable, "-c", "from pyresample_b200 import _cabi; print(_cabi.build(force=True))"], env=env)
