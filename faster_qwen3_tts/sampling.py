"""Alias of `qwen3_tts_cuda_graphs_b200.sampling`: the reference's callers and tests import `faster_qwen3_tts.sampling` by name."""
from qwen3_tts_cuda_graphs_b200.sampling import *  # noqa: F401,F403
from qwen3_tts_cuda_graphs_b200 import sampling as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
