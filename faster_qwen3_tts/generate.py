"""Alias of `qwen3_tts_cuda_graphs_b200.generate`: the reference's callers and tests import `faster_qwen3_tts.generate` by name."""
from qwen3_tts_cuda_graphs_b200.generate import *  # noqa: F401,F403
from qwen3_tts_cuda_graphs_b200 import generate as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
