"""inference_gateway_b200 -- B200-native streaming-response hot path of inference-gateway v0.24.0.

The product is libssegpu.so (csrc/: hand-written CUDA for sm_100a behind the C ABI of include/sse_gpu.h).
This package only loads it, drives it and generates synthetic workloads. It never imports oracle/.
"""
from . import _abi  # noqa: F401
from .engine import BatchResult, SseEngine  # noqa: F401

__all__ = ["SseEngine", "BatchResult"]
