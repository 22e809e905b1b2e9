import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# GPU tests whose kernels changed after their last hardware run (re-verified under the CPU emulation only) run after the ones whose kernels did not, and the
# never-run ones (test_zz_*) last, so that `pytest -x` reaches every hardware-verified test first.
_LATE = {"test_filter_gpu.py": 1, "test_trajectory_gpu.py": 1, "test_verify_filters_gpu.py": 1, "test_zz_sift_prune_gpu.py": 2, "test_zz_sift_detect_gpu.py": 3}


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: _LATE.get(os.path.basename(str(it.fspath)), 0))          # stable: the order inside each group is unchanged


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
