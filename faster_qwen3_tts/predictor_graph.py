"""Alias of `qwen3_tts_cuda_graphs_b200.predictor_graph`: the reference's callers and tests import `faster_qwen3_tts.predictor_graph` by name."""
from qwen3_tts_cuda_graphs_b200.predictor_graph import *  # noqa: F401,F403
from qwen3_tts_cuda_graphs_b200 import predictor_graph as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
