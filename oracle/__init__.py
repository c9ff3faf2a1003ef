"""CPU oracle for the pyresample kd-tree neighbour resampling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyresample_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` do.  It is a numpy (+ scipy cKDTree,
+ a small C brute-force checker) restatement of the reference algorithm; every
function cites the reference ``file:line`` it follows (paths relative to the
pyresample source tree).

Parity status: PINNED.  ``tests/test_oracle_golden.py`` replays the reference's
own known-answer tests and golden grids (pyresample/test/test_kd_tree.py,
test_swath.py, test_spatial_mp.py, test_data_reduce.py, test_geometry/test_area.py)
through this oracle.  pykdtree and PROJ are third-party dependencies that are NOT
in the reference tree (setup.py:27-28, unpinned); their behaviour is restated from
their published algorithms and anchored on those reference tests.  What no
reference test pins (tie-breaking, float32 rounding, eps>0) is stated in DESIGN.md.
"""
