"""ORACLE (test infrastructure only — nothing under qwen3_tts_cuda_graphs_b200/ may import this): the reference's streaming
codec-decode policy, `faster_qwen3_tts/model.py:737-826` (twins at `:941-1001` and `:1106-1166`), restated on plain Python / numpy
around an arbitrary `decode(codes[T, Q]) -> 1-D samples` callable.

What the reference does per yielded chunk of codec frames:

* until calibrated (`model.py:773-802`): decode EVERYTHING generated so far — with the ICL reference codes in front when there are
  any (`:777-780`), whose share of the waveform is cut off proportionally, `int(ref_len / total_len * len(audio))` (`:790-795`) —
  and hand out the samples beyond what was handed out before (`:799-800`); once at least `max(25, chunk_size)` frames exist
  (`:741-742`, `:802`), remember samples-per-frame = generated samples / generated frames (`:803`);
* afterwards (`:804-823`): decode the last `25 + n_new` frames (fewer context frames at the very start, `:806`) and drop
  `int(round(n_ctx * samples_per_frame))` samples from the front (`:819-821`).

Pinned: pure control flow over an injected decoder, no arithmetic of the absent `qwen_tts` involved — the reference lines above ARE
the specification; `tests/test_stream_policy_cpu.py` holds `model.WindowedDecode` against this restatement sample for sample."""
from typing import Callable, Iterable, Iterator, Optional

import numpy as np

CONTEXT_FRAMES = 25  # model.py:741


def stream_decode_policy(chunks: Iterable[np.ndarray], decode: Callable[[np.ndarray], np.ndarray],
                         ref_codes: Optional[np.ndarray], chunk_size: int) -> Iterator[np.ndarray]:
    """chunks: int arrays [n_i, Q] as `fast_generate_streaming` yields them; yields the new samples per chunk."""
    need = max(CONTEXT_FRAMES, chunk_size)     # frames before samples-per-frame is trusted
    history = np.zeros((0, 0), dtype=np.int64)
    handed_out = 0                             # samples of generated audio already yielded (calibration phase)
    per_frame = None
    for chunk in chunks:
        chunk = np.asarray(chunk)
        history = chunk.copy() if history.size == 0 else np.concatenate([history, chunk], axis=0)
        total, fresh = history.shape[0], chunk.shape[0]
        if per_frame is None:
            if ref_codes is None:
                generated = np.asarray(decode(history)).reshape(-1)
            else:
                both = np.concatenate([np.asarray(ref_codes), history], axis=0)
                wav = np.asarray(decode(both)).reshape(-1)
                generated = wav[int(ref_codes.shape[0] / max(both.shape[0], 1) * len(wav)):]
            out = generated[handed_out:]
            handed_out = len(generated)
            if total >= need:
                per_frame = len(generated) / total
        else:
            first = max(0, total - fresh - CONTEXT_FRAMES)
            wav = np.asarray(decode(history[first:])).reshape(-1)
            context = (total - first) - fresh
            out = wav[int(round(context * per_frame)):] if context > 0 else wav
        yield out


def full_decode_policy(codes: np.ndarray, decode: Callable[[np.ndarray], np.ndarray], ref_codes: Optional[np.ndarray]) -> np.ndarray:
    """Non-streaming decode, `model.py:634-656`: ICL reference codes go in front of the generated ones (`:636-641`), one decode
    (`:642`), and the reference's share of the waveform — `int(ref_len / total_len * len(audio))` samples — is cut off (`:645-655`)."""
    if ref_codes is None:
        return np.asarray(decode(np.asarray(codes))).reshape(-1)
    both = np.concatenate([np.asarray(ref_codes), np.asarray(codes)], axis=0)
    wav = np.asarray(decode(both)).reshape(-1)
    return wav[int(ref_codes.shape[0] / max(both.shape[0], 1) * len(wav)):]
