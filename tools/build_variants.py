"""Kernel experiments: build variants of the library (macro settings) next to the product's, units float_8_3 and
float_16_0 only (the C3 / C4 search kernels)."""
import os
import subprocess
import sys

sys.path.insert(0, ".")
VARIANTS = {
    "v1": "-DPRS_HOME_FACTOR=1.3f",
    "v2": "-DPRS_HOME_FACTOR=1.5f",
    "v3": "-DPRS_HOME_FACTOR=1.3f -DPRS_HOME_MAX=5",
    "v4": "-DPRS_HOME_FACTOR=1.5f -DPRS_HOME_MAX=6",
    "v5": "-DPRS_HOME_FACTOR=1.8f -DPRS_HOME_MAX=6",
}
for name, flags in VARIANTS.items():
    env = dict(os.environ, PRS_B200_LIBRARY=os.path.abspath("pyresample_b200/libprs_%s.so" % name), PRS_NVCC_FLAGS=flags,
               PRS_BUILD_ONLY="float_8_3,float_16_0")
    subprocess.check_call([sys.executable, "-c", "from pyresample_b200 import _cabi; print(_cabi.build(force=True))"], env=env)
