"""Checkpoint directory -> TTSConfig + weight dicts ("real-checkpoint loader (safetensors -> packed arena)").

What the reference does at this point is one call into the un-vendored `qwen_tts` package,
`Qwen3TTSModel.from_pretrained(model_name, device_map=..., torch_dtype=...)` (`faster_qwen3_tts/model.py:107-112`), after which
it only dereferences module attributes (`model.py:114-119`, `predictor_graph.py:52-57`).  Here the same directory layout —

    <dir>/config.json                      top-level ids + "talker_config" { ..., "code_predictor_config": {...} } (+ "speaker_encoder_config")
    <dir>/model.safetensors | model-0000x-of-0000y.safetensors + model.safetensors.index.json
    <dir>/speech_tokenizer/config.json     "decoder_config" {...} (+ "encoder_config")
    <dir>/speech_tokenizer/model.safetensors
    <dir>/tokenizer.json | vocab.json + merges.txt

— is read straight into the tensors `weights.pack_arena` and `codec.CodecDecoder` consume: the state-dict keys are the module
paths the reference itself names (`talker.model.layers.N.self_attn.q_proj.weight`, `talker.code_predictor.lm_head.i.weight`, ...),
so no per-tensor translation table is needed for the LM; the few places where the upstream tree is only *recalled*
(SURVEY.md Appendix B — codebooks stored Mimi-style as `embedding_sum / cluster_usage`, 1x1 convs stored as [out, in, 1]) go
through the alias rules at the bottom of this file, which are data, not code.  Every tensor is shape-checked against
`weights.tensor_specs` / `codec.codec_tensor_specs`; a missing or mis-shaped tensor raises with its name (no silent random init).

No checkpoint exists offline: the loader is exercised end to end on directories written by `export_checkpoint` (the inverse
mapping, also what `tests/test_checkpoint_cpu.py` round-trips bit for bit) — parity against a real upstream file is unpinned.
"""
from __future__ import annotations

import dataclasses
import json
import os
import re
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

from .config import CodecDecoderConfig, PredictorConfig, TalkerConfig, TTSConfig, preset


class CheckpointError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------
# config.json
# ------------------------------------------------------------------------------------------------
def _pick(cls, src: dict, base, **force):
    """Fields of dataclass `cls` taken from `src` where present, else from the instance `base`."""
    kw = {}
    for f in dataclasses.fields(cls):
        if f.name in force:
            kw[f.name] = force[f.name]
        elif f.name in src and src[f.name] is not None:
            v = src[f.name]
            kw[f.name] = tuple(v) if isinstance(getattr(base, f.name), tuple) else v
        else:
            kw[f.name] = getattr(base, f.name)
    return cls(**kw)


def config_from_hf(raw: dict, codec_raw: Optional[dict] = None) -> TTSConfig:
    """HF `config.json` (+ the speech tokenizer's) -> TTSConfig.  Unknown keys are ignored, absent ones keep the preset value
    for the model size (`tts_model_size`: "0b6" | "1b7"), so a partial config still loads and a full one overrides everything."""
    tc = dict(raw.get("talker_config") or {})
    pc = dict(tc.pop("code_predictor_config", None) or raw.get("code_predictor_config") or {})
    size = str(raw.get("tts_model_size") or ("1b7" if tc.get("hidden_size", 1024) >= 2048 else "0b6")).lower()
    kind = str(raw.get("tts_model_type") or "base").lower()
    # defaults for absent keys come from the preset of the nearest named size; the size tag itself is kept as written
    name = {"0b6": "0.6B", "1b7": "1.7B"}.get(size) or ("1.7B" if tc.get("hidden_size", 1024) >= 2048 else "0.6B")
    kinds = {"base": "Base", "custom_voice": "CustomVoice", "voice_design": "VoiceDesign"}
    base = preset(f"{name}-{kinds.get(kind, 'Base')}")
    for k in ("codec_language_id", "spk_id"):
        if isinstance(tc.get(k), dict):
            tc[k] = {str(a).lower(): int(b) for a, b in tc[k].items()}
    if isinstance(tc.get("spk_is_dialect"), dict):
        tc["spk_is_dialect"] = {str(a).lower(): b for a, b in tc["spk_is_dialect"].items()}
    rope = tc.get("rope_parameters") or {}
    if "rope_theta" not in tc and "rope_theta" in rope:
        tc["rope_theta"] = rope["rope_theta"]
    talker = _pick(TalkerConfig, tc, base.talker)
    if "num_code_groups" not in pc:
        pc["num_code_groups"] = talker.num_code_groups
    predictor = _pick(PredictorConfig, pc, base.predictor)
    cd = dict((codec_raw or {}).get("decoder_config") or codec_raw or {})
    if "output_sample_rate" in (codec_raw or {}):
        cd.setdefault("sample_rate", codec_raw["output_sample_rate"])
    codec = _pick(CodecDecoderConfig, cd, base.codec)
    top = {f.name for f in dataclasses.fields(TTSConfig)} - {"talker", "predictor", "codec"}
    kw = {k: raw[k] for k in top if k in raw and raw[k] is not None}
    kw.setdefault("tts_model_type", kind)
    kw.setdefault("tts_model_size", size)
    return dataclasses.replace(base, talker=talker, predictor=predictor, codec=codec, **kw)


def config_to_hf(cfg: TTSConfig) -> Tuple[dict, dict]:
    """Inverse of config_from_hf: (config.json, speech_tokenizer/config.json)."""
    d = cfg.to_dict()
    talker, predictor, codec = d.pop("talker"), d.pop("predictor"), d.pop("codec")
    talker["code_predictor_config"] = predictor
    top = dict(d)
    top.update(architectures=["Qwen3TTSForConditionalGeneration"], model_type="qwen3_tts", talker_config=talker)
    codec = {k: (list(v) if isinstance(v, tuple) else v) for k, v in codec.items()}
    return top, {"model_type": "qwen3_tts_tokenizer_12hz", "output_sample_rate": codec["sample_rate"], "decoder_config": codec}


def read_config(path: str) -> TTSConfig:
    with open(os.path.join(path, "config.json")) as f:
        raw = json.load(f)
    codec_raw = None
    cpath = os.path.join(path, "speech_tokenizer", "config.json")
    if os.path.exists(cpath):
        with open(cpath) as f:
            codec_raw = json.load(f)
    return config_from_hf(raw, codec_raw)


# ------------------------------------------------------------------------------------------------
# safetensors
# ------------------------------------------------------------------------------------------------
class TensorDir:
    """Name -> tensor over every *.safetensors file of a directory (sharded or not), opened lazily: a tensor is read from disk
    when asked for, so host memory holds one matrix at a time while the arena is packed."""

    def __init__(self, path: str):
        from safetensors import safe_open

        self._open = safe_open
        self.path = path
        self.where: Dict[str, str] = {}
        idx = os.path.join(path, "model.safetensors.index.json")
        if os.path.exists(idx):
            with open(idx) as f:
                self.where = {k: os.path.join(path, v) for k, v in json.load(f)["weight_map"].items()}
        else:
            files = sorted(fn for fn in os.listdir(path) if fn.endswith(".safetensors")) if os.path.isdir(path) else []
            if not files:
                raise CheckpointError(f"{path}: no *.safetensors file")
            for fn in files:
                with safe_open(os.path.join(path, fn), framework="pt") as f:
                    for k in f.keys():
                        self.where[k] = os.path.join(path, fn)
        self._handles: Dict[str, object] = {}

    def __contains__(self, name: str) -> bool:
        return name in self.where

    def keys(self) -> Iterable[str]:
        return self.where.keys()

    def get(self, name: str) -> torch.Tensor:
        fn = self.where[name]
        h = self._handles.get(fn)
        if h is None:
            h = self._handles[fn] = self._open(fn, framework="pt")
        return h.get_tensor(name)

    def close(self):
        self._handles.clear()


def _resolve(td: TensorDir, name: str, prefixes: Tuple[str, ...], rules) -> torch.Tensor:
    for pre in prefixes:
        if pre + name in td:
            return td.get(pre + name)
    for pat, fn in rules:
        m = re.fullmatch(pat, name)
        if m:
            for pre in prefixes:
                t = fn(td, pre, m)
                if t is not None:
                    return t
    raise CheckpointError(f"{td.path}: tensor {name!r} not found (tried prefixes {list(prefixes)} and {len(rules)} alias rules)")


def _checked(name: str, t: torch.Tensor, shape: tuple) -> torch.Tensor:
    if tuple(t.shape) != tuple(shape):
        if t.numel() == int(torch.tensor(shape).prod()) and tuple(s for s in t.shape if s != 1) == tuple(s for s in shape if s != 1):
            return t.reshape(shape)  # 1x1 convs stored as [out, in, 1]
        raise CheckpointError(f"tensor {name!r} has shape {tuple(t.shape)}, the config implies {tuple(shape)}")
    return t


def load_lm_weights(path: str, cfg: TTSConfig, skip: Callable[[str], bool] = lambda n: False) -> Dict[str, torch.Tensor]:
    """bf16 CPU tensors keyed as `weights.tensor_specs` (what `pack_arena` and the oracle take)."""
    from .weights import tensor_specs

    td = TensorDir(path)
    out = {}
    try:
        for name, shape, _ in tensor_specs(cfg):
            if skip(name):
                continue
            out[name] = _checked(name, _resolve(td, name, LM_PREFIXES, LM_RULES), shape).to(torch.bfloat16).contiguous()
    finally:
        td.close()
    return out


def load_codec_weights(path: str, ccfg: CodecDecoderConfig) -> Dict[str, torch.Tensor]:
    """fp32 CPU tensors keyed as `codec.codec_tensor_specs` (CodecDecoder rounds matrices to bf16 itself)."""
    from .codec import codec_tensor_specs

    td = TensorDir(path)
    out = {}
    try:
        k = 3
        for pre in CODEC_PREFIXES:
            if pre + "pre_conv.conv.weight" in td:
                k = td.get(pre + "pre_conv.conv.weight").shape[-1]
        for name, shape, _ in codec_tensor_specs(ccfg, pre_conv_kernel=k):
            out[name] = _checked(name, _resolve(td, name, CODEC_PREFIXES, CODEC_RULES), shape).to(torch.float32).contiguous()
    finally:
        td.close()
    return out


def load_prefixed(path: str, prefixes: Tuple[str, ...]) -> Dict[str, torch.Tensor]:
    """Every tensor under one of `prefixes`, prefix stripped (speaker encoder, codec encoder)."""
    td = TensorDir(path)
    out = {}
    try:
        for k in td.keys():
            for pre in prefixes:
                if k.startswith(pre):
                    out[k[len(pre):]] = td.get(k)
                    break
    finally:
        td.close()
    return out


# ------------------------------------------------------------------------------------------------
# export (tests, and `python -m qwen3_tts_cuda_graphs_b200.checkpoint --export` for a synthetic stand-in on disk)
# ------------------------------------------------------------------------------------------------
def export_checkpoint(path: str, cfg: TTSConfig, lm: Dict[str, torch.Tensor], codec: Optional[Dict[str, torch.Tensor]] = None,
                      extra: Optional[Dict[str, torch.Tensor]] = None, codec_extra: Optional[Dict[str, torch.Tensor]] = None,
                      shards: int = 1, mimi_codebooks: bool = False, extra_config: Optional[dict] = None,
                      codec_extra_config: Optional[dict] = None) -> None:
    """Write the directory layout of the module docstring.  `shards` > 1 writes an index file; `mimi_codebooks` stores the
    codec codebooks the way Mimi-style quantisers do (embedding_sum / cluster_usage) to exercise the alias rules."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    top, ctop = config_to_hf(cfg)
    top.update(extra_config or {})
    ctop.update(codec_extra_config or {})
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(top, f, indent=1)
    tensors = {k: v.contiguous() for k, v in lm.items()}
    tensors.update({k: v.contiguous() for k, v in (extra or {}).items()})
    names = list(tensors)
    if shards <= 1:
        save_file(tensors, os.path.join(path, "model.safetensors"))
    else:
        wm = {}
        for s in range(shards):
            fn = f"model-{s + 1:05d}-of-{shards:05d}.safetensors"
            part = {k: tensors[k] for k in names[s::shards]}
            save_file(part, os.path.join(path, fn))
            wm.update({k: fn for k in part})
        with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {}, "weight_map": wm}, f)
    if codec is not None:
        sdir = os.path.join(path, "speech_tokenizer")
        os.makedirs(sdir, exist_ok=True)
        with open(os.path.join(sdir, "config.json"), "w") as f:
            json.dump(ctop, f, indent=1)
        ct = {}
        for k, v in codec.items():
            m = re.fullmatch(r"quantizer\.codebook\.(\d+)", k)
            if m and mimi_codebooks:
                g = int(m.group(1))
                nsem = cfg.codec.num_semantic_quantizers
                stem = f"quantizer.rvq_first.vq.layers.{g}" if g < nsem else f"quantizer.rvq_rest.vq.layers.{g - nsem}"
                usage = torch.full((v.shape[0],), 2.0)
                ct[f"decoder.{stem}._codebook.embedding_sum"] = (v.float() * 2.0).contiguous()
                ct[f"decoder.{stem}._codebook.cluster_usage"] = usage
            elif k.endswith("output_proj.weight") and mimi_codebooks:
                ct["decoder." + k] = v.unsqueeze(-1).contiguous()
            else:
                ct["decoder." + k] = v.contiguous()
        ct.update({k: v.contiguous() for k, v in (codec_extra or {}).items()})
        save_file(ct, os.path.join(sdir, "model.safetensors"))


# ------------------------------------------------------------------------------------------------
# alias rules (recalled upstream layout; unverifiable offline — extend here, not in code)
# ------------------------------------------------------------------------------------------------
LM_PREFIXES = ("", "model.")
CODEC_PREFIXES = ("decoder.", "", "model.decoder.")


def _mimi_codebook(td: TensorDir, pre: str, m) -> Optional[torch.Tensor]:
    g = int(m.group(1))
    for nsem in (1,):
        stem = f"{pre}quantizer.rvq_first.vq.layers.{g}" if g < nsem else f"{pre}quantizer.rvq_rest.vq.layers.{g - nsem}"
        if f"{stem}._codebook.embedding_sum" in td:
            usage = td.get(f"{stem}._codebook.cluster_usage").float().clamp(min=1e-5)
            return td.get(f"{stem}._codebook.embedding_sum").float() / usage[:, None]
        if f"{stem}._codebook.embed" in td:
            return td.get(f"{stem}._codebook.embed")
    return None


LM_RULES: List[tuple] = []
CODEC_RULES: List[tuple] = [(r"quantizer\.codebook\.(\d+)", _mimi_codebook)]


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="write a synthetic (random-init) checkpoint directory in the layout the loader reads")
    ap.add_argument("--export", required=True, help="output directory")
    ap.add_argument("--preset", default="tiny")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    from .codec import init_codec_synthetic
    from .weights import init_synthetic

    c = preset(a.preset)
    export_checkpoint(a.export, c, init_synthetic(c, seed=a.seed), init_codec_synthetic(c.codec, seed=a.seed + 1))
    print(f"wrote {a.export}")
