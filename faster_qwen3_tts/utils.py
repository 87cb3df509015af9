"""`faster_qwen3_tts.utils` of the reference holds one helper, a context manager that silences the flash-attn import warning of the
upstream `qwen_tts` package (`utils.py`, used at `model.py:101-102`).  Nothing here imports `qwen_tts` or flash-attn, so the helper is
kept for callers that import it and does nothing."""
import contextlib


@contextlib.contextmanager
def suppress_flash_attn_warning():
    yield
