import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Native artefacts (product library, oracle, reference tools) built once per session."""
    import helpers
    helpers.ensure_built()
    return helpers

# the host stream packer reads this knob once per process: the tests run with its whole-line non-temporal path on (aligned
# outputs) so that it stays covered; unaligned outputs in test_host_stream_packer cover the default 16-byte path
os.environ.setdefault("FM_HOSTPACK_LINE", "1")
