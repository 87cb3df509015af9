"""GPU tests of the public FasterQwen3TTS API (reference: faster_qwen3_tts/model.py; test strategy of
tests/test_e2e_parity.py — structural validity, streaming == non-streaming, instruct-prefix invariance, error
behaviour) on the tiny synthetic preset."""
import os
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TEXT = "Short parity test."  # tests/test_e2e_parity.py:175


@pytest.fixture(scope="module")
def ref_wav(tmp_path_factory):
    p = tmp_path_factory.mktemp("audio") / "ref.wav"
    sr = 24000
    t = np.arange(int(1.2 * sr)) / sr
    pcm = (0.3 * np.sin(2 * np.pi * 220 * t) * 32767).astype(np.int16)
    with wave.open(str(p), "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes(pcm.tobytes())
    return str(p)


def load(name, **kw):
    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS

    return FasterQwen3TTS.from_pretrained(name, device="cuda", dtype=torch.bfloat16, max_seq_len=256, **kw)


@pytest.fixture(scope="module")
def base():
    m = load("tiny-Base")
    yield m
    m.model.engine.close()


def _valid_codes(codes, cfg):
    """_assert_codec_output_valid of tests/test_e2e_parity.py:40-101."""
    assert codes.dim() == 2 and codes.shape[1] == 16
    assert (codes >= 0).all()
    assert (codes[:, 0] < cfg.vocab_size - 1024).all() or (codes[:, 0] != cfg.codec_eos_token_id).all()
    assert (codes[:, 0] != cfg.codec_eos_token_id).all()


def test_rejects_non_cuda_device():
    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS

    with pytest.raises(ValueError):
        FasterQwen3TTS.from_pretrained("tiny-Base", device="cpu")


def test_generate_is_not_implemented(base):
    with pytest.raises(NotImplementedError):
        base.generate("hello")


def test_voice_clone_returns_float32_audio(base, ref_wav):
    audio, sr = base.generate_voice_clone(TEXT, "English", ref_wav, "", max_new_tokens=12, do_sample=False)
    assert sr == 24000 == base.sample_rate
    assert isinstance(audio, list) and len(audio) == 1
    a = audio[0]
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 1
    assert a.size == base.model.model.speech_tokenizer.decoder.n_samples(12)
    assert np.isfinite(a).all() and np.abs(a).max() <= 1.0
    assert base._warmed_up and len(base._voice_prompt_cache) == 1


def test_streaming_codes_equal_non_streaming(base, ref_wav):
    """tests/test_e2e_parity.py:726-780 with chunk_size=8, greedy and sampled (same seed)."""
    from qwen3_tts_cuda_graphs_b200.generate import fast_generate
    from qwen3_tts_cuda_graphs_b200.streaming import fast_generate_streaming

    for do_sample in (False, True):
        m, talker, config, tie, tam, tth, tpe, _ = base._prepare_generation(TEXT, ref_wav, "", language="English",
                                                                            non_streaming_mode=True)
        kw = dict(talker=talker, talker_input_embeds=tie, attention_mask=tam, trailing_text_hiddens=tth, tts_pad_embed=tpe,
                  config=config, predictor_graph=base.predictor_graph, talker_graph=base.talker_graph, max_new_tokens=21,
                  do_sample=do_sample, seed=1234)
        full, timing = fast_generate(**kw)
        chunks = list(fast_generate_streaming(chunk_size=8, **kw))
        assert [c.shape[0] for c, _ in chunks] == [8, 8, 5]
        assert torch.equal(torch.cat([c for c, _ in chunks]), full)
        assert chunks[0][1]["prefill_ms"] > 0 and chunks[-1][1]["is_final"]
        assert set(timing) == {"prefill_ms", "decode_s", "steps", "ms_per_step", "steps_per_s"}
        _valid_codes(full, config)


def test_streaming_audio_matches_full_decode(base, ref_wav):
    """Hybrid accumulate -> sliding-window policy (model.py:737-826): chunk audio concatenates to the full decode
    (exact while accumulating; the windowed phase re-decodes with 25 frames of left context)."""
    base.predictor_graph.do_sample = False  # like tests/test_e2e_parity.py:208-215: predictor greedy => same codes twice
    kw = dict(max_new_tokens=40, do_sample=False)
    full, sr = base.generate_voice_clone(TEXT, "English", ref_wav, "", **kw)
    chunks = list(base.generate_voice_clone_streaming(TEXT, "English", ref_wav, "", chunk_size=8, **kw))
    assert len(chunks) == 5 and all(c[1] == sr for c in chunks)
    cat = np.concatenate([c[0] for c in chunks])
    n_acc = sum(len(c[0]) for c in chunks[:4])  # 32 frames >= 25: phase 1 covers the first four chunks
    assert np.array_equal(cat[:n_acc], full[0][:n_acc])
    assert abs(len(cat) - len(full[0])) <= 1920
    assert chunks[0][2]["chunk_index"] == 0 and chunks[-1][2]["total_steps_so_far"] == 40
    base.predictor_graph.do_sample = True


def test_icl_mode_prepends_ref_codes_and_trims(base, ref_wav):
    audio, sr = base.generate_voice_clone(TEXT, "English", ref_wav, "reference words", max_new_tokens=10, do_sample=False,
                                          xvec_only=False)
    dec = base.model.model.speech_tokenizer.decoder
    n_ref = int(round((1.2 + 0.5) * 12.5))
    total = dec.n_samples(n_ref + 10)
    assert len(audio[0]) == total - int(n_ref / (n_ref + 10) * total)


def test_icl_non_streaming_decode_skips_the_reference_part_bit_exactly(base, ref_wav, monkeypatch):
    """_decode_full (model.py:634-656) cuts the reference clip's samples off; the vocoder does not compute them (skip_samples) and the
    kept ones do not change by a bit."""
    base.predictor_graph.do_sample = False
    try:
        kw = dict(max_new_tokens=14, do_sample=False, xvec_only=False)
        fast, _ = base.generate_voice_clone(TEXT, "English", ref_wav, "reference words", **kw)
        monkeypatch.setenv("FQ3C_TAIL_ONLY", "0")
        full, _ = base.generate_voice_clone(TEXT, "English", ref_wav, "reference words", **kw)
    finally:
        base.predictor_graph.do_sample = True
    assert fast[0].shape == full[0].shape == (14 * 1920,) and np.array_equal(fast[0], full[0])


def test_icl_streaming_skips_the_reference_part_without_changing_a_sample(base, ref_wav, monkeypatch):
    """While the window policy still accumulates (model.py:737-826) an ICL stream decodes reference + generated frames and throws
    the reference part away; with the causal length law that cut is known up front and the vocoder only computes what the kept
    samples can see.  Same samples as with the skip turned off (FQ3C_TAIL_ONLY=0), chunk by chunk."""
    kw = dict(max_new_tokens=40, do_sample=False, xvec_only=False, chunk_size=8)
    base.predictor_graph.do_sample = False
    try:
        fast = [a for a, _, _ in base.generate_voice_clone_streaming(TEXT, "English", ref_wav, "reference words", **kw)]
        dec = base.model.model.speech_tokenizer.decoder
        n_ref = int(round((1.2 + 0.5) * 12.5))
        assert any(isinstance(k, tuple) and k[0] == n_ref + 8 and k[1] > 0 for k in dec._plans), "the first ICL decode did not skip the reference part"
        monkeypatch.setenv("FQ3C_TAIL_ONLY", "0")
        full = [a for a, _, _ in base.generate_voice_clone_streaming(TEXT, "English", ref_wav, "reference words", **kw)]
    finally:
        base.predictor_graph.do_sample = True
    assert [len(a) for a in fast] == [len(a) for a in full] == [8 * 1920] * 5
    for a, b in zip(fast, full):
        assert np.array_equal(a, b)


def test_prefill_longer_than_cache_raises(base, ref_wav):
    with pytest.raises(RuntimeError, match="Input is too long"):
        base.generate_voice_clone(" ".join(["word"] * 300), "English", ref_wav, "", max_new_tokens=4)


def test_model_type_guards(base):
    with pytest.raises(ValueError):
        base.generate_custom_voice(TEXT, "aiden", "English")
    with pytest.raises(ValueError):
        base.generate_voice_design(TEXT, "a calm voice", "English")


def test_custom_voice_and_instruct_prefix():
    m = load("tiny-CustomVoice")
    try:
        with pytest.raises(ValueError):
            m.generate_custom_voice(TEXT, "nobody", "English")
        audio, sr = m.generate_custom_voice(TEXT, "aiden", "English", max_new_tokens=6, do_sample=False)
        assert len(audio[0]) == m.model.model.speech_tokenizer.decoder.n_samples(6)
        chunks = list(m.generate_custom_voice_streaming(TEXT, "aiden", "English", max_new_tokens=6, do_sample=False, chunk_size=4))
        assert [c[2]["chunk_steps"] for c in chunks] == [4, 2]
        # instruct prepends exactly instruct_len rows and leaves the suffix bit-identical (test_e2e_parity.py:1020-1049)
        _, _, _, tie0, tam0, _, _ = m._prepare_generation_custom(TEXT, "English", "aiden", None)
        _, _, _, tie1, tam1, _, _ = m._prepare_generation_custom(TEXT, "English", "aiden", "speak slowly")
        n = tie1.shape[1] - tie0.shape[1]
        assert n == len(m.model.tokenizer.instruct("speak slowly"))
        assert torch.equal(tie1[:, n:], tie0)
        # dialect speakers override the language id (model.py:387-393)
        _, _, _, tie_d, _, _, _ = m._prepare_generation_custom(TEXT, "Chinese", "dylan", None)
        _, _, _, tie_c, _, _, _ = m._prepare_generation_custom(TEXT, "Chinese", "aiden", None)
        assert not torch.equal(tie_d[:, 5], tie_c[:, 5])
    finally:
        m.model.engine.close()


def test_voice_design():
    m = load("tiny-VoiceDesign")
    try:
        a0, _ = m.generate_voice_design(TEXT, "a deep calm voice", "English", max_new_tokens=6, do_sample=False)
        a1, _ = m.generate_voice_design(TEXT, "a bright fast voice", "English", max_new_tokens=6, do_sample=False)
        assert len(a0[0]) == len(a1[0]) and not np.array_equal(a0[0], a1[0])  # instruct changes output (:1052-1080)
        chunks = list(m.generate_voice_design_streaming(TEXT, "a deep calm voice", "English", max_new_tokens=6,
                                                        do_sample=False, chunk_size=3))
        assert len(chunks) == 2
    finally:
        m.model.engine.close()


def test_voice_clone_batch_matches_single_requests(ref_wav):
    """Request-parallel decode (BASELINE configs[4]): greedy, three texts through a 2-stream engine (one full group and one
    partial) must give, per text, exactly the waveform the bs = 1 call gives."""
    m = load("tiny-Base", max_streams=2)
    m.predictor_graph.do_sample = False  # the sub-talker policy is the graph's own mutable state (predictor_graph.py:34-60)
    try:
        texts = [TEXT, "A second, somewhat longer sentence for the batch.", "Third."]
        kw = dict(max_new_tokens=10, do_sample=False, repetition_penalty=1.0)
        singles = [m.generate_voice_clone(t, "English", ref_wav, "", **kw)[0][0] for t in texts]
        batch, sr = m.generate_voice_clone_batch(texts, "English", ref_wav, "", **kw)
        assert sr == m.sample_rate and len(batch) == len(texts)
        for a, b in zip(singles, batch):
            assert a.shape == b.shape and np.array_equal(a, b)
    finally:
        m.model.engine.close()


def test_streaming_with_stateful_codec_equals_non_streaming_audio(base, ref_wav, monkeypatch):
    """FQ3_STATEFUL_CODEC=1: the chunks of generate_voice_clone_streaming concatenate to EXACTLY the audio of the non-streaming
    call (same codes: greedy) — the stateful codec stream reproduces the full decode bit for bit (split-K off on both sides),
    where the reference's windowed policy only approximates it."""
    monkeypatch.setenv("FQ3C_SPLITK", "0")
    base.predictor_graph.do_sample = False
    kw = dict(max_new_tokens=27, min_new_tokens=27, do_sample=False)   # 27 frames: a plan no other test has cached
    full, sr = base.generate_voice_clone(TEXT, "English", ref_wav, "", **kw)
    monkeypatch.setenv("FQ3_STATEFUL_CODEC", "1")
    chunks = [a for a, _, _ in base.generate_voice_clone_streaming(TEXT, "English", ref_wav, "", chunk_size=8, **kw)]
    assert [len(c) for c in chunks] == [8 * 1920, 8 * 1920, 8 * 1920, 3 * 1920]
    got = np.concatenate(chunks)
    assert got.shape == full[0].shape
    assert np.array_equal(got, full[0])


def test_reference_example_flow_with_a_saved_speaker_embedding(base, ref_wav, tmp_path):
    """examples/extract_speaker.py + examples/generate_with_embedding.py line by line: x-vector saved to a .pt file, loaded back
    into a voice_clone_prompt dict, prompt built with `_build_talker_inputs_local`, `fast_generate` called positionally,
    `speech_tokenizer.decode` with the list form — same codes and audio as `generate_voice_clone(xvec_only=True)`."""
    from qwen3_tts_cuda_graphs_b200.generate import fast_generate

    items = base.model.create_voice_clone_prompt(ref_audio=ref_wav, ref_text="", x_vector_only_mode=True)  # extract_speaker.py:32-38
    path = str(tmp_path / "speaker.pt")
    torch.save(items[0].ref_spk_embedding.cpu(), path)
    spk_emb = torch.load(path, weights_only=True).to("cuda:0")                                              # generate_with_embedding.py:29
    vcp = dict(ref_code=[None], ref_spk_embedding=[spk_emb], x_vector_only_mode=[True], icl_mode=[False])
    input_ids = base.model._tokenize_texts([base.model._build_assistant_text(TEXT)])
    tie, tam, tth, tpe = base._build_talker_inputs_local(
        m=base.model.model, input_ids=input_ids, ref_ids=[None], voice_clone_prompt=vcp, languages=["English"], speakers=None,
        non_streaming_mode=False)
    base._warmup(tie.shape[1])
    talker, config = base.model.model.talker, base.model.model.config.talker_config
    base.predictor_graph.do_sample = False
    try:
        talker.rope_deltas = None
        codec_ids, timing = fast_generate(talker, tie, tam, tth, tpe, config, base.predictor_graph, base.talker_graph,
                                          temperature=0.9, top_k=50, do_sample=False, max_new_tokens=20)
        assert codec_ids.shape == (20, 16) and timing["steps"] == 20 and timing["ms_per_step"] > 0
        wavs, sr = base.model.speech_tokenizer.decode([{"audio_codes": codec_ids.to(base.device)}])
        assert sr == 24000 and len(wavs) == 1
        audio, _ = base.generate_voice_clone(TEXT, "English", ref_wav, "", xvec_only=True, non_streaming_mode=False,
                                             do_sample=False, max_new_tokens=20)
        assert np.array_equal(wavs[0].flatten().float().cpu().numpy(), audio[0])
    finally:
        base.predictor_graph.do_sample = True
