import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)        # test helpers (wavutil)


try:
    # Property tests (hypothesis) run DERANDOMIZED in the suite: the same examples on every machine, no example database.
    # Explore with `pytest --hypothesis-seed=N -p no:cacheprovider` (any failure found that way becomes a fixed regression case).
    from hypothesis import settings as _hyp_settings
    _hyp_settings.register_profile("suite", derandomize=True, database=None, deadline=None)
    _hyp_settings.load_profile("suite")
except ImportError:     # hypothesis is optional: its tests skip themselves
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
