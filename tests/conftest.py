import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _gpu_available():
    """True when the product library loads and sees a device.  J2P_EXPECT_GPU=1 (set on the GPU
    boxes by tools/run_gpu_checks.sh) turns "no device" into a hard failure instead of a skip, so a
    broken box or a missing .so cannot pass as 'all skipped'."""
    try:
        from jpeg2png_b200 import abi
        return abi.load_product().j2p_device_count() > 0
    except Exception:       # noqa: BLE001 — library not built
        return False


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker('gpu')]
    if not gpu_items:
        return
    selected = config.getoption('-m') or ''
    if _gpu_available():
        return
    if os.environ.get('J2P_EXPECT_GPU') == '1' or selected.strip() == 'gpu':
        return              # `-m gpu` or an explicit expectation: let the tests fail loudly
    skip = pytest.mark.skip(reason='no CUDA device / libjpeg2png_b200.so not built (GPU tests run with -m gpu on the B200 box)')
    for it in gpu_items:
        it.add_marker(skip)
