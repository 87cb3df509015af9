"""PredictorGraph — the 15-codebook code-predictor operator (`faster_qwen3_tts/predictor_graph.py`) on fq3.

`run(pred_input[1,2,H]) -> LongTensor[15]` executes the whole `_full_loop` (predictor_graph.py:115-167) —
small_to_mtp, 2-token prefill, 15 heads, 14 embed+decode steps, in-kernel sampling — as ONE persistent
kernel launch whose KV state (<= 17 positions) never leaves L2 and whose weights are streamed with an
evict-last hint.  The sampling policy attributes are mutable and read at run time; the reference freezes
them at capture time (model.py:124-133), which is the default here too because nobody mutates them after
warm-up (tests mutate them before, tests/test_e2e_parity.py:208-215).
"""
from __future__ import annotations

import torch

from .engine import Engine, SubPolicy


class PredictorGraph:
    def __init__(self, engine: Engine, stream_idx: int = 0, do_sample: bool = True, top_k: int = 50, top_p: float = 1.0,
                 temperature: float = 0.9, seed: int = 0):
        if stream_idx != 0:
            raise ValueError("the PredictorGraph operator seam addresses stream 0 only (predictor_graph.py:70-71 is bs = 1 as well)")
        self.engine = engine
        self.stream_idx = stream_idx
        self.device = engine.device
        self.dtype = torch.bfloat16
        pc = engine.cfg.predictor
        self.num_layers = pc.num_hidden_layers
        self.hidden_size = pc.hidden_size
        self.num_code_groups = pc.num_code_groups
        self.num_codebooks = pc.num_codebooks
        self.max_seq = 2 + self.num_codebooks
        self.do_sample, self.top_k, self.top_p, self.temperature = do_sample, top_k, top_p, temperature
        self.seed = seed
        self._calls = 0
        self.captured = False
        self.graph = None
        self.last_logits = None

    def policy(self) -> SubPolicy:
        return SubPolicy(self.do_sample, self.top_k, self.top_p, self.temperature)

    @torch.inference_mode()
    def capture(self, num_warmup: int = 3):
        """predictor_graph.py:169-202 — nothing to capture."""
        self.captured = True

    @torch.inference_mode()
    def run(self, pred_input: torch.Tensor, want_logits: bool = False) -> torch.Tensor:
        """predictor_graph.py:204-214."""
        self._calls += 1
        codes, logits = self.engine.predictor_run(
            self.stream_idx, pred_input, self.policy(), seed=self.seed + 0x51ED * self._calls, want_logits=want_logits
        )
        self.last_logits = logits
        return codes
