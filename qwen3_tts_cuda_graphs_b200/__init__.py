"""B200-native Qwen3-TTS decode engine behind the FasterQwen3TTS API (drop-in for the hot path of
andimarafioti/qwen3-tts-cuda-graphs).  See DESIGN.md."""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "FasterQwen3TTS":
        from .model import FasterQwen3TTS

        return FasterQwen3TTS
    raise AttributeError(name)


__all__ = ["FasterQwen3TTS"]
