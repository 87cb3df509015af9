"""The algorithmic byte model behind `roofline.achieved` (`weights.param_bytes`, bench.py) against SURVEY.md §8(d)'s per-frame figures
and against an independent count of the tensors `weights.tensor_specs` declares for the arena — so the roofline's numerator cannot
drift from the architecture the engine actually streams."""
import pytest

from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.weights import param_bytes, tensor_specs

MB = 1e6


def test_byte_model_equals_survey_8d_for_0p6b():
    pb = param_bytes(preset("0.6B-Base"))
    assert pb["talker_step"] / MB == pytest.approx(887.2, abs=0.1)        # (440.47 M + 3.15 M head) x 2 B
    assert pb["predictor_pass"] / MB == pytest.approx(157.3, abs=0.1)     # 78.65 M layer params x 2 B
    assert pb["predictor_heads"] / MB == pytest.approx(62.9, abs=0.1)     # 15 x 2048 x 1024 x 2 B
    assert pb["frame_streaming"] / MB == pytest.approx(3309.7, abs=0.2)   # 887.2 + 15 x 157.3 + 62.9: what `achieved` is computed from
    assert pb["frame_read_once"] / MB == pytest.approx(1107.4, abs=0.2)
    assert pb["frame_streaming"] == pb["talker_step"] + 15 * pb["predictor_pass"] + pb["predictor_heads"]


def test_byte_model_equals_survey_8d_for_1p7b():
    pb = param_bytes(preset("1.7B-Base"))
    assert pb["talker_step"] / MB == pytest.approx(2831, abs=1)           # (1409.4 M + 6.29 M) x 2 B
    s2m = 2 * (1024 * 2048 + 1024)                                        # small_to_mtp Linear 2048 -> 1024 with bias: 4.2 MB per pass
    assert pb["predictor_pass"] / MB == pytest.approx(157.3 + s2m / MB, abs=0.1)


@pytest.mark.parametrize("name", ["0.6B-Base", "1.7B-CustomVoice", "tiny-Base"])
def test_byte_model_equals_the_tensors_the_arena_declares(name):
    """Independent count: sum the element counts of the declared tensors by name instead of using the closed-form layer formula."""
    cfg = preset(name)
    specs = tensor_specs(cfg)
    items = specs.items() if isinstance(specs, dict) else specs

    def numel(shape):
        n = 1
        for d in shape:
            n *= int(d)
        return n

    sizes = {}
    for entry in items:
        key, shape = entry[0], entry[1]
        sizes[key] = numel(shape)
    talker = sum(n for k, n in sizes.items() if k.startswith("talker.model.layers.") or k == "talker.model.norm.weight")
    head = sizes["talker.codec_head.weight"]
    pred = sum(n for k, n in sizes.items()
               if k.startswith("talker.code_predictor.model.layers.") or k == "talker.code_predictor.model.norm.weight")
    s2m = sum(n for k, n in sizes.items() if k.startswith("talker.code_predictor.small_to_mtp_projection."))
    heads = sum(n for k, n in sizes.items() if k.startswith("talker.code_predictor.lm_head."))
    pb = param_bytes(cfg)
    assert pb["talker_step"] == 2 * (talker + head)
    assert pb["predictor_pass"] == 2 * (pred + s2m)
    assert pb["predictor_heads"] == 2 * heads
