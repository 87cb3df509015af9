"""Checkpoint directory -> config / weights / tokenizer (reference: faster_qwen3_tts/model.py:107-119, the
`Qwen3TTSModel.from_pretrained` call whose module tree the reference dereferences).  No real checkpoint exists offline: the
loader is held against directories written by its own inverse (`export_checkpoint`), bit for bit."""
import json
import os

import pytest
import torch

from helpers import write_tiny_checkpoint
from qwen3_tts_cuda_graphs_b200 import checkpoint as ck
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.weights import pack_arena


@pytest.mark.parametrize("name", ["0.6B-Base", "1.7B-Base", "0.6B-CustomVoice", "1.7B-CustomVoice", "1.7B-VoiceDesign"])
def test_config_round_trip(name):
    cfg = preset(name)
    top, codec = ck.config_to_hf(cfg)
    assert ck.config_from_hf(json.loads(json.dumps(top)), json.loads(json.dumps(codec))) == cfg


def test_partial_config_keeps_preset_values():
    raw = {"tts_model_type": "custom_voice", "tts_model_size": "1b7",
           "talker_config": {"hidden_size": 2048, "spk_id": {"Aiden": 3100}, "code_predictor_config": {"num_hidden_layers": 5}}}
    cfg = ck.config_from_hf(raw)
    ref = preset("1.7B-CustomVoice")
    assert cfg.talker.intermediate_size == ref.talker.intermediate_size and cfg.talker.spk_id == {"aiden": 3100}
    assert cfg.predictor == ref.predictor and cfg.tts_model_type == "custom_voice"


@pytest.mark.parametrize("shards,mimi", [(1, False), (3, True)])
def test_directory_round_trip_is_bit_exact(tmp_path, shards, mimi):
    cfg0 = preset("tiny-Base")
    # the tiny preset is not "0b6"/"1b7": write its dims through a config whose size tag is 0b6 and every field explicit
    d = str(tmp_path / "ckpt")
    cfg, lm, codec, _, _ = write_tiny_checkpoint(d, cfg=cfg0.__class__(**{**cfg0.__dict__, "tts_model_size": "0b6"}), shards=shards,
                                                 mimi_codebooks=mimi)
    got_cfg = ck.read_config(d)
    assert got_cfg == cfg
    got = ck.load_lm_weights(d, got_cfg)
    assert set(got) == set(lm)
    for k in lm:
        assert got[k].dtype == torch.bfloat16 and torch.equal(got[k], lm[k]), k
    a, b = pack_arena(cfg, lm, 64, "cpu"), pack_arena(got_cfg, got, 64, "cpu")
    assert a.offsets == b.offsets and torch.equal(a.buf, b.buf)
    gc = ck.load_codec_weights(os.path.join(d, "speech_tokenizer"), got_cfg.codec)
    assert set(gc) == set(codec)
    for k in codec:
        assert torch.equal(gc[k], codec[k].float()), k


def test_missing_and_misshaped_tensors_raise_with_their_name(tmp_path):
    from safetensors.torch import load_file, save_file

    cfg0 = preset("tiny-Base")
    d = str(tmp_path / "ckpt")
    cfg, lm, _, _, _ = write_tiny_checkpoint(d, cfg=cfg0.__class__(**{**cfg0.__dict__, "tts_model_size": "0b6"}), with_frontend=False)
    f = os.path.join(d, "model.safetensors")
    t = {k: v.clone() for k, v in load_file(f).items()}  # the file is rewritten below: do not keep views of its mapping
    victim = "talker.model.layers.1.mlp.down_proj.weight"
    keep = t.pop(victim)
    save_file(t, f)
    with pytest.raises(ck.CheckpointError, match="down_proj"):
        ck.load_lm_weights(d, cfg)
    t[victim] = keep[:, :-1].contiguous()
    save_file(t, f)
    with pytest.raises(ck.CheckpointError, match="shape"):
        ck.load_lm_weights(d, cfg)
    with pytest.raises(ck.CheckpointError, match="safetensors"):
        ck.TensorDir(str(tmp_path))


def test_tokenizer_templates_have_the_layout_the_prompt_builder_slices(tmp_path):
    """model.py:435,454,466,480,509 slice [:, :3], [:, 3:-5], [:, 3:-2]: 3 role ids, text, 5 / 2 trailer ids."""
    from qwen3_tts_cuda_graphs_b200.base_model import HFTokenizer

    cfg0 = preset("tiny-Base")
    d = str(tmp_path / "ckpt")
    cfg, *_ = write_tiny_checkpoint(d, cfg=cfg0.__class__(**{**cfg0.__dict__, "tts_model_size": "0b6"}), with_frontend=False)
    tok = HFTokenizer(d)
    text = tok.encode("Hello world, the quick brown fox!")
    assert len(text) == 8 and 15 not in text
    role = [cfg.im_start_token_id, cfg.assistant_token_id, cfg.newline_token_id]
    a = tok.assistant("Hello world, the quick brown fox!")
    assert a[:3] == role and a[3:-5] == text and a[-5:] == [cfg.im_end_token_id, cfg.newline_token_id] + role
    r = tok.ref("reference clip says this.")
    assert r[:3] == role and r[-2:] == [cfg.im_end_token_id, cfg.newline_token_id] and len(r) == 3 + 5 + 2
    i = tok.instruct("speak slowly")
    assert i[:3] == [cfg.im_start_token_id, cfg.user_token_id, cfg.newline_token_id] and len(i) == 3 + 2 + 2
