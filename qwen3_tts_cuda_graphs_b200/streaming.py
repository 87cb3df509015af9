"""fast_generate_streaming — chunked generation (`faster_qwen3_tts/streaming.py:19-188`) on the fq3 engine.

Yields `(codec_chunk int64 [<=chunk_size, 16], timing)` with the reference's timing keys
(streaming.py:162-169).  In the fused path each chunk is ONE persistent-kernel launch of `chunk_size`
frames and one host synchronisation — the same frames, in the same order, as the non-streaming call
(the Philox draw counter lives in the stream state, so sampling is chunk-size invariant too).
"""
from __future__ import annotations

import os
import time
from typing import Generator, Optional, Tuple

import torch

from .engine import SamplingPolicy
from .generate import _fresh_seed, fused_prefill, is_fused, operator_frames


@torch.inference_mode()
def fast_generate_streaming(
    talker,
    talker_input_embeds: torch.Tensor,
    attention_mask: torch.Tensor,
    trailing_text_hiddens: torch.Tensor,
    tts_pad_embed: torch.Tensor,
    config,
    predictor_graph,
    talker_graph,
    max_new_tokens: int = 2048,
    min_new_tokens: int = 2,
    temperature: float = 0.9,
    top_k: int = 50,
    top_p: float = 1.0,
    do_sample: bool = True,
    repetition_penalty: float = 1.05,
    chunk_size: int = 12,
    seed: Optional[int] = None,
    launch_events: Optional[list] = None,
) -> Generator[Tuple[torch.Tensor, dict], None, None]:
    dev = talker_input_embeds.device
    if is_fused(talker, predictor_graph, talker_graph):
        policy = SamplingPolicy(
            do_sample=do_sample, top_k=top_k, top_p=top_p, temperature=temperature,
            repetition_penalty=repetition_penalty, min_new_tokens=min_new_tokens, suppress_tail=1024,
            seed=_fresh_seed() if seed is None else seed,
        )
        sub = predictor_graph.policy()
        t0 = time.time()
        eng, idx = fused_prefill(talker_graph, predictor_graph, talker_input_embeds, attention_mask,
                                 trailing_text_hiddens, tts_pad_embed, policy)
        torch.cuda.synchronize()
        t_prefill = time.time() - t0
        budget = min(max_new_tokens, eng.max_frames)
        emitted, chunk_idx = 0, 0
        # Overlapped streaming (FQ3_OVERLAP_CODEC=0 turns it off): from the second chunk on, the frame loop runs on the engine's
        # reduced grid (128 of 148 CTAs: the frame time is flat down to there) and the next chunk is launched BEFORE the current
        # one is yielded, so the caller's codec decode (on a side stream, model.py) runs beside it on the free SMs.  The first
        # chunk keeps the full grid and the plain order: time to first audio must not wait behind a speculative launch.
        reduced = eng.reduced_grid() if hasattr(eng, "reduced_grid") else 0
        overlap = os.environ.get("FQ3_OVERLAP_CODEC", "1") != "0" and reduced > 0
        in_flight = 0  # frames of a chunk that is already running
        if reduced > 0:
            eng.set_decode_grid(0)  # first chunk (and a stream left behind by an aborted generator): full grid

        def launch(n):
            if launch_events is not None:  # bench.py: CUDA-event bracket of the persistent-kernel launch alone
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            eng.decode_frames(1, n, policy, sub)
            if launch_events is not None:
                e1.record()
                launch_events.append((e0, e1, n))

        while emitted < budget:
            t1 = time.time()
            if in_flight == 0:
                launch(min(chunk_size, budget - emitted))
            in_flight = 0
            st = eng.status(idx)
            dt = time.time() - t1
            n_new = st.n_frames - emitted
            if n_new > 0:
                chunk = eng.read_codes(idx, emitted, n_new).to(dev)
                emitted += n_new
                info = {
                    "chunk_index": chunk_idx, "chunk_steps": n_new, "prefill_ms": t_prefill * 1000 if chunk_idx == 0 else 0,
                    "decode_ms": dt * 1000, "total_steps_so_far": emitted, "is_final": n_new < chunk_size,
                }
                chunk_idx += 1
                if overlap and chunk_idx >= 2 and n_new == chunk_size and not st.done and emitted < budget:
                    eng.set_decode_grid(reduced)
                    ready = torch.cuda.Event()
                    ready.record()                       # the codes of this chunk are complete here ...
                    info["codes_ready"] = ready          # ... the side stream waits for this, not for the next chunk
                    in_flight = min(chunk_size, budget - emitted)
                    launch(in_flight)
                # like the reference, only a trailing partial chunk is flagged final (streaming.py:175-188)
                try:
                    yield chunk, info
                except GeneratorExit:
                    if in_flight:
                        eng.status(idx)                  # the consumer left: let the speculative chunk finish
                    if overlap:
                        eng.set_decode_grid(0)
                    raise
                if n_new < chunk_size:
                    break
            if st.done or n_new == 0:
                break
        if overlap:
            eng.set_decode_grid(0)                       # other callers of the engine get the full grid back
        return

    tinfo: dict = {}
    buf, total, chunk_idx = [], 0, 0
    t_chunk = time.time()
    for frame in operator_frames(talker, talker_input_embeds, attention_mask, trailing_text_hiddens, tts_pad_embed,
                                 config, predictor_graph, talker_graph, max_new_tokens, min_new_tokens, temperature,
                                 top_k, top_p, do_sample, repetition_penalty, seed, tinfo):
        buf.append(frame)
        if len(buf) >= chunk_size:
            torch.cuda.synchronize()
            total += len(buf)
            yield torch.stack(buf), {
                "chunk_index": chunk_idx, "chunk_steps": len(buf),
                "prefill_ms": tinfo.get("prefill_s", 0.0) * 1000 if chunk_idx == 0 else 0,
                "decode_ms": (time.time() - t_chunk) * 1000, "total_steps_so_far": total, "is_final": False,
            }
            buf, chunk_idx, t_chunk = [], chunk_idx + 1, time.time()
    if buf:
        torch.cuda.synchronize()
        total += len(buf)
        yield torch.stack(buf), {
            "chunk_index": chunk_idx, "chunk_steps": len(buf),
            "prefill_ms": tinfo.get("prefill_s", 0.0) * 1000 if chunk_idx == 0 else 0,
            "decode_ms": (time.time() - t_chunk) * 1000, "total_steps_so_far": total, "is_final": True,
        }
