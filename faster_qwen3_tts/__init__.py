"""Drop-in alias: `from faster_qwen3_tts import FasterQwen3TTS` resolves to the B200-native implementation.

Mirrors the reference package's public names (`faster_qwen3_tts/__init__.py:4-7`) so callers written against
andimarafioti/qwen3-tts-cuda-graphs (cli.py, demo/server.py, examples/*, benchmarks/*) import unchanged.
"""
from qwen3_tts_cuda_graphs_b200.model import FasterQwen3TTS  # noqa: F401
from qwen3_tts_cuda_graphs_b200.predictor_graph import PredictorGraph  # noqa: F401
from qwen3_tts_cuda_graphs_b200.talker_graph import TalkerGraph  # noqa: F401
from qwen3_tts_cuda_graphs_b200.generate import fast_generate  # noqa: F401
from qwen3_tts_cuda_graphs_b200.streaming import fast_generate_streaming  # noqa: F401
from qwen3_tts_cuda_graphs_b200.sampling import apply_repetition_penalty, sample_logits  # noqa: F401

__version__ = "0.2.4"  # the reference release whose API surface this package mirrors (faster_qwen3_tts/__init__.py:6)
__all__ = ["FasterQwen3TTS", "PredictorGraph", "TalkerGraph", "fast_generate", "fast_generate_streaming",
           "sample_logits", "apply_repetition_penalty"]
