"""The reference's own known-answer tests (tests/reference_cases.py) replayed through the CUDA product."""
import pytest

import reference_cases as rc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", rc.ALL_CASES, ids=lambda f: f.__name__)
def test_reference_case_through_cuda(cuda_lib, case):
    from pyresample_b200 import kd_tree
    case(kd_tree)
