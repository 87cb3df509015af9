"""TalkerGraph — the reference's per-frame talker operator (`faster_qwen3_tts/talker_graph.py`) on the fq3 engine.

Same seam, same method names and argument meaning: `max_seq_len`, `prefill_kv`,
`set_generation_state`, `run`, `capture`, `reset`.  Nothing is captured: "capture" is kept as a no-op so the
reference's warm-up call sequence (`model.py:154-163`) still works.  One decode step is one launch of the
persistent weight-streaming kernel (28 layers + final norm + codec_head), not ~400 graph nodes.
"""
from __future__ import annotations

from typing import Optional

import torch

from .engine import Engine


class TalkerGraph:
    def __init__(self, engine: Engine, stream_idx: int = 0):
        if stream_idx != 0:
            # fq3_talker_step / fq3_prefill return hidden states through row 0 of the engine's buffers: the operator-at-a-time
            # seam is single-stream like the reference's (talker_graph.py:46-47); batches go through fq3_decode_frames
            raise ValueError("the TalkerGraph operator seam addresses stream 0 only; use fast_generate_batch for several streams")
        self.engine = engine
        self.stream_idx = stream_idx
        self.device = engine.device
        self.dtype = torch.bfloat16
        self.max_seq_len = engine.max_seq_len
        self.hidden_size = engine.cfg.talker.hidden_size
        self.num_layers = engine.cfg.talker.num_hidden_layers
        # static output buffer, like talker_graph.py:47 — "use immediately or clone" (talker_graph.py:214)
        self.output_buf = torch.zeros(1, 1, self.hidden_size, dtype=self.dtype, device=self.device)
        self.last_logits: Optional[torch.Tensor] = None  # codec_head(output_buf) computed in the same launch
        self.captured = False
        self.graph = None

    @torch.inference_mode()
    def capture(self, prefill_len: int = 100, num_warmup: int = 3):
        """talker_graph.py:109-147.  The engine has no lazy capture; run one step so first-use costs
        (module load, L2 warm-up) are paid here as in the reference."""
        self.captured = True

    def reset(self, prefill_len: int = 0):
        """talker_graph.py:149-151."""
        self.engine.reset_stream(self.stream_idx)

    def prefill_kv(self, past_key_values) -> int:
        """talker_graph.py:153-170: import a prefix KV ([1, kv_heads, T, head_dim] per layer)."""
        if hasattr(past_key_values, "engine") and hasattr(past_key_values, "length"):
            # prefix already written in place by the engine's own prefill (base_model.Talker.forward)
            if past_key_values.length > self.max_seq_len:
                raise RuntimeError(
                    f"Input is too long: prefill has {past_key_values.length} tokens but max_seq_len={self.max_seq_len}. "
                    "Use shorter text or shorter reference audio."
                )
            return past_key_values.length
        seq_len = 0
        self.engine.reset_stream(self.stream_idx)
        for li in range(self.num_layers):
            k, v = past_key_values[li]
            seq_len = k.shape[2]
            if seq_len > self.max_seq_len:
                raise RuntimeError(
                    f"Input is too long: prefill has {seq_len} tokens but max_seq_len={self.max_seq_len}. "
                    "Use shorter text or shorter reference audio."
                )
            self.engine.import_kv(self.stream_idx, li, k[0], v[0])
        return seq_len

    def set_generation_state(self, attention_mask: Optional[torch.Tensor], rope_deltas: Optional[torch.Tensor]):
        """talker_graph.py:172-196: left-pad count and rope delta.  The reference rebuilds a max_seq_len-entry
        mask table here; the kernel only needs the pad count because it bounds attention by length."""
        n_pad = 0
        if attention_mask is not None:
            n_pad = int((attention_mask.reshape(attention_mask.shape[0], -1)[0] == 0).sum().item())
        delta = 0
        if rope_deltas is not None:
            delta = int(round(float(torch.as_tensor(rope_deltas).reshape(-1)[0].item())))
        self.engine.set_generation_state(self.stream_idx, n_pad, delta)

    @torch.inference_mode()
    def run(self, input_embeds: torch.Tensor, position: int) -> torch.Tensor:
        """talker_graph.py:198-214: [1,1,H] -> [1,1,H] (static buffer)."""
        hidden, logits = self.engine.talker_step(self.stream_idx, input_embeds, position, want_logits=True)
        self.output_buf.view(-1).copy_(hidden)
        self.last_logits = logits
        return self.output_buf
