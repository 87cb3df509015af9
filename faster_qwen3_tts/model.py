"""Alias of `qwen3_tts_cuda_graphs_b200.model`: the reference's callers and tests import `faster_qwen3_tts.model` by name."""
from qwen3_tts_cuda_graphs_b200.model import *  # noqa: F401,F403
from qwen3_tts_cuda_graphs_b200 import model as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
