import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", params=["fused", "automaton", "copy", "split", "v1"])
def engine(request):
    """One GPU context per kernel generation (the product path: libssegpu.so through its C ABI).
    split = produce / decode / finalize pipeline (default, zero-copy frames); copy = the same with SSE_FLAG_COPY_OUT; v2 = fused producer-consumer kernel; v1 = first generation."""
    from inference_gateway_b200 import SseEngine
    eng = SseEngine(device=0, max_conns=4096, bytes_per_batch=8 << 20, carry_slot_bytes=32768, n_slots=2,
                    flags={"v1": 1, "split": 4, "fused": 0, "automaton": 16, "copy": 8}[request.param])
    yield eng
    eng.close()
