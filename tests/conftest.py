import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", params=["split", "templates", "copy", "fused", "fused_templates"])
def engine(request):
    """One GPU context per kernel generation (the product path: libssegpu.so through its C ABI).
    split = produce / sort / decode / finalize pipeline (default, zero-copy frames); templates = the same with skeleton-template replay; copy = default with SSE_FLAG_COPY_OUT; fused / fused_templates = the single-pass tile kernel."""
    from inference_gateway_b200 import SseEngine
    eng = SseEngine(device=0, max_conns=4096, bytes_per_batch=8 << 20, carry_slot_bytes=32768, n_slots=2,
                    flags={"split": 0, "fused": 4, "fused_templates": 4 | 16, "templates": 16, "copy": 8}[request.param])
    yield eng
    eng.close()
