"""GPU check of the OPT-IN fused cloth / body pair (hmSDF_Tets.split(fused=True)).  The very last file of
the suite: the path is not the default, was developed on the kernel emulation and has never run on a GPU, so a failure
here must not keep the tests of the default path (run with -x) from running.  Body: tests/test_z_configs.py."""
import pytest
import torch

from tests import test_z_configs as Z

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True, params=["sort", "static", "scan"])
def edges_mode(request):
    """Every test runs three times: on the general path (per-call radix sort + run-length scan of the crossing-edge keys),
    on the static edge table path (tet stream + bitmap over the grid's sorted edge list, built once per tet array) and on
    the edge-scan path (walk over the static edge list instead of the tet stream; the default of a training run)."""
    from d3human_code_b200 import extract as E
    E.set_static_edges("0" if request.param == "sort" else "1")
    E.set_edge_scan(request.param == "scan")
    yield request.param
    E.set_static_edges("auto")
    E.set_edge_scan(True)


@pytest.mark.parametrize("res,field", [(24, "capsule"), (12, "adv")])
def test_fused_pair_equals_two_calls(dev, res, field):
    Z.fused_pair_equals_two_calls(dev, res, field)
