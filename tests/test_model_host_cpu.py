"""Host logic of `FasterQwen3TTS` around the engine (reference `faster_qwen3_tts/model.py:202-330, 843-850,
1017-1018`) against doubles of the base model — no GPU: the voice-prompt cache and its key, the x-vector and ICL branches of
`_prepare_generation` (0.5 s of silence appended to the reference clip, reference-text ids, reference codes handed on only in ICL
mode), the one-time warm-up, `rope_deltas` reset per request, the model-type guards and the 0.6B instruct rule."""
import types
import wave

import numpy as np
import pytest
import torch

from qwen3_tts_cuda_graphs_b200.model import FasterQwen3TTS


class BaseDouble:
    def __init__(self, kind="base", size="0b6"):
        talker = types.SimpleNamespace(rope_deltas="stale")
        self.model = types.SimpleNamespace(talker=talker, config=types.SimpleNamespace(talker_config="talker-config"),
                                           speech_tokenizer=types.SimpleNamespace(sample_rate=24000),
                                           tts_model_type=kind, tts_model_size=size)
        self.frontend = None
        self.prompt_calls = []
        self.validated = []

    def _build_assistant_text(self, t):
        return f"<assistant>{t}"

    def _build_ref_text(self, t):
        return f"<ref>{t}"

    def _build_instruct_text(self, t):
        return f"<instruct>{t}"

    def _tokenize_texts(self, texts):
        return [torch.tensor([[len(t)] + [ord(c) for c in t[:3]]]) for t in texts]

    def create_voice_clone_prompt(self, ref_audio, ref_text="", x_vector_only_mode=False):
        self.prompt_calls.append((ref_audio, ref_text, x_vector_only_mode))
        return [types.SimpleNamespace(ref_code=None if x_vector_only_mode else torch.zeros(7, 16, dtype=torch.long),
                                      ref_spk_embedding=torch.ones(4), ref_text=ref_text, x_vector_only_mode=x_vector_only_mode,
                                      icl_mode=not x_vector_only_mode)]

    def _prompt_items_to_voice_clone_prompt(self, items):
        return dict(ref_code=[i.ref_code for i in items], ref_spk_embedding=[i.ref_spk_embedding for i in items],
                    x_vector_only_mode=[i.x_vector_only_mode for i in items], icl_mode=[i.icl_mode for i in items])

    def _validate_languages(self, langs):
        self.validated.append(("lang", tuple(langs)))

    def _validate_speakers(self, spk):
        self.validated.append(("spk", tuple(spk)))


class Graph:
    def __init__(self):
        self.captures = []

    def capture(self, **kw):
        self.captures.append(kw)


@pytest.fixture()
def tts(monkeypatch):
    base = BaseDouble()
    t = FasterQwen3TTS(base, Graph(), Graph(), device="cuda", dtype=torch.bfloat16, max_seq_len=2048)
    built = []

    def build(m, input_ids, ref_ids, voice_clone_prompt, languages, speakers, non_streaming_mode, instruct_ids=None):
        built.append(dict(input_ids=input_ids, ref_ids=ref_ids, vcp=voice_clone_prompt, languages=languages, speakers=speakers,
                          nsm=non_streaming_mode, instruct_ids=instruct_ids))
        return torch.zeros(1, 21, 8), torch.ones(1, 21, dtype=torch.long), torch.zeros(1, 1, 8), torch.zeros(1, 1, 8)

    monkeypatch.setattr(t, "_build_talker_inputs_local", build)
    t.built = built
    return t


@pytest.fixture()
def clip(tmp_path):
    p = tmp_path / "voice.wav"
    with wave.open(str(p), "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(16000)
        wf.writeframes((np.sin(np.arange(8000) / 10.0) * 8000).astype("<i2").tobytes())
    return str(p)


def test_xvector_branch_caches_the_voice_prompt_per_key(tts, clip):
    base = tts.model
    assert tts.sample_rate == 24000 and tts._warmed_up is False and tts._voice_prompt_cache == {}
    m, talker, cfg, tie, tam, tth, tpe, ref_codes = tts._prepare_generation("Hello", clip, "ignored in x-vector mode", language="English")
    assert base.prompt_calls == [(clip, "", True)]                     # model.py:234-238: x-vector only, no transcript
    assert ref_codes is None and cfg == "talker-config" and talker.rope_deltas is None     # model.py:285
    b = tts.built[-1]
    assert b["vcp"]["x_vector_only_mode"] == [True] and b["vcp"]["icl_mode"] == [False] and b["vcp"]["ref_code"] == [None]
    assert b["ref_ids"] == [None] and b["languages"] == ["English"] and b["speakers"] is None and b["instruct_ids"] == [None]
    assert tts._warmed_up and tts.talker_graph.captures == [dict(prefill_len=21, num_warmup=3)]    # model.py:154-163, 278-279
    talker.rope_deltas = "stale again"
    tts._prepare_generation("Another line", clip, "ignored in x-vector mode", language=None)
    assert len(base.prompt_calls) == 1 and len(tts._voice_prompt_cache) == 1                       # cache hit (model.py:230-231)
    assert tts.built[-1]["languages"] == ["Auto"] and talker.rope_deltas is None
    assert len(tts.talker_graph.captures) == 1                                                     # warmed up once
    tts._prepare_generation("Hello", clip, "other transcript", language="English")                 # the key holds ref_text too
    assert len(base.prompt_calls) == 2 and len(tts._voice_prompt_cache) == 2


def test_icl_branch_appends_silence_and_hands_the_reference_codes_on(tts, clip):
    base = tts.model
    *_, ref_codes = tts._prepare_generation("Hello", clip, "what the clip says", language="English", xvec_only=False,
                                            non_streaming_mode=True, instruct="speak slowly")
    (audio, sr), ref_text, xvec = base.prompt_calls[-1]
    assert sr == 16000 and len(audio) == 8000 + 8000 and np.all(audio[8000:] == 0) and audio.dtype == np.float32   # 0.5 s of silence
    assert ref_text == "what the clip says" and xvec is False
    assert ref_codes.shape == (7, 16)                                                              # prepended to the codec input later
    b = tts.built[-1]
    assert b["nsm"] is True and b["vcp"]["icl_mode"] == [True]
    assert torch.equal(b["ref_ids"][0], base._tokenize_texts(["<ref>what the clip says"])[0])      # model.py:258-261
    assert torch.equal(b["instruct_ids"][0], base._tokenize_texts(["<instruct>speak slowly"])[0])
    tts._prepare_generation("Hello", clip, "what the clip says", language="English", xvec_only=False, append_silence=False)
    (audio, _), _, _ = base.prompt_calls[-1]
    assert len(audio) == 8000 and len(tts._voice_prompt_cache) == 2                                # append_silence is part of the key


def test_model_type_guards_and_the_small_model_instruct_rule(tts):
    with pytest.raises(NotImplementedError):
        tts.generate("hello")                                                                      # model.py:165-183
    with pytest.raises(ValueError, match="does not support custom voice"):
        tts.generate_custom_voice("x", "aiden", "English")                                         # model.py:843-844
    with pytest.raises(ValueError, match="does not support voice design"):
        tts.generate_voice_design("x", "a calm voice", "English")                                  # model.py:1017-1018
    with pytest.raises(ValueError, match="does not support custom voice"):
        next(tts.generate_custom_voice_streaming("x", "aiden", "English"))
    with pytest.raises(NotImplementedError, match="parity_mode"):
        next(tts.generate_voice_clone_streaming("x", "English", "v.wav", "", parity_mode=True))
    tts.model.model.tts_model_type = "custom_voice"
    tts._check_custom("English", "aiden")
    assert tts.model.validated == [("lang", ("English",)), ("spk", ("aiden",))]
    m, talker, cfg, *_ = tts._prepare_generation_custom("Hi", "English", "aiden", instruct="whisper")
    b = tts.built[-1]
    assert b["speakers"] == ["aiden"] and b["vcp"] is None and b["nsm"] is False and b["instruct_ids"][0] is not None
    assert talker.rope_deltas is None
