"""The drop-in import surface: every module path the reference's callers and tests import (`faster_qwen3_tts`, `.model`, `.generate`,
`.streaming`, `.sampling`, `.talker_graph`, `.predictor_graph`, `.utils`, `.cli`) resolves to this package, and the reference's own
CPU-runnable test of the public class (`tests/test_sample_rate.py:10-29`: sample-rate inference and its fallbacks) passes against it."""
import types


def test_reference_module_paths_resolve_to_this_package():
    import faster_qwen3_tts
    from faster_qwen3_tts.cli import build_parser
    from faster_qwen3_tts.generate import fast_generate
    from faster_qwen3_tts.model import FasterQwen3TTS
    from faster_qwen3_tts.predictor_graph import PredictorGraph
    from faster_qwen3_tts.sampling import apply_repetition_penalty, sample_logits
    from faster_qwen3_tts.streaming import fast_generate_streaming
    from faster_qwen3_tts.talker_graph import TalkerGraph
    from faster_qwen3_tts.utils import suppress_flash_attn_warning
    import qwen3_tts_cuda_graphs_b200 as impl

    assert FasterQwen3TTS is faster_qwen3_tts.FasterQwen3TTS is impl.FasterQwen3TTS
    assert TalkerGraph is faster_qwen3_tts.TalkerGraph and PredictorGraph is faster_qwen3_tts.PredictorGraph
    assert fast_generate is faster_qwen3_tts.fast_generate and fast_generate_streaming is faster_qwen3_tts.fast_generate_streaming
    assert sample_logits is faster_qwen3_tts.sample_logits and apply_repetition_penalty is faster_qwen3_tts.apply_repetition_penalty
    assert set(faster_qwen3_tts.__all__) >= {"FasterQwen3TTS"}  # faster_qwen3_tts/__init__.py:4-7
    with suppress_flash_attn_warning():
        assert {a.dest for a in build_parser()._actions} >= {"device", "dtype"}


def _dummy_graph():
    return object()


def test_uses_speech_tokenizer_sample_rate_when_available():
    from faster_qwen3_tts.model import FasterQwen3TTS

    base_model = types.SimpleNamespace(model=types.SimpleNamespace(speech_tokenizer=types.SimpleNamespace(sample_rate=24000)))
    assert FasterQwen3TTS(base_model, _dummy_graph(), _dummy_graph()).sample_rate == 24000


def test_falls_back_to_base_model_sample_rate():
    from faster_qwen3_tts.model import FasterQwen3TTS

    assert FasterQwen3TTS(types.SimpleNamespace(sample_rate=22050), _dummy_graph(), _dummy_graph()).sample_rate == 22050


def test_defaults_to_24khz_when_sample_rate_unavailable():
    from faster_qwen3_tts.model import FasterQwen3TTS

    assert FasterQwen3TTS(types.SimpleNamespace(model=types.SimpleNamespace()), _dummy_graph(), _dummy_graph()).sample_rate == 24000


def test_constructor_attributes_of_the_reference():
    """model.py:39-47: the attributes callers read off the object."""
    from faster_qwen3_tts.model import FasterQwen3TTS

    m = FasterQwen3TTS(types.SimpleNamespace(sample_rate=24000), "p", "t", device="cuda", max_seq_len=512)
    assert (m.predictor_graph, m.talker_graph, m.device, m.max_seq_len, m._warmed_up, m._voice_prompt_cache) == ("p", "t", "cuda", 512, False, {})
