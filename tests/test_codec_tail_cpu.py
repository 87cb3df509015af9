"""Receptive-field walk of the tail-only streaming decode (codec.vocoder_tail_starts) against a brute-force dependency
closure over the same layer graph: every row a kept sample can see must lie at or after the op's first row, and the first
rows must not be wastefully early (they are tight up to the floor of the transposed conv)."""
import pytest

from qwen3_tts_cuda_graphs_b200.codec import VOCODER_DILATIONS, vocoder_tail_starts


def brute_force(rows0, rates, skip, trim="both"):
    """Needed-row sets, walking the layers backwards with explicit per-row dependencies."""
    right = trim == "right"
    lvl = [rows0]
    for r in rates:
        lvl.append(lvl[-1] * r if right else (lvl[-1] - 1) * r)
    n = lvl[-1]
    need = set(range(skip, n))                      # final conv output rows
    fin = min(need)
    need = {r - t for r in need for t in range(7) if r - t >= 0}   # its input (block output) rows
    units = [[None] * 3 for _ in rates]
    tconv = [None] * len(rates)
    for i in range(len(rates) - 1, -1, -1):
        for j in (2, 1, 0):
            units[i][j] = min(need)                 # rows of x_out / hs this unit must produce
            d = VOCODER_DILATIONS[j]
            need = need | {r - t * d for r in need for t in range(7) if r - t * d >= 0}   # conv1 input (and the residual: same rows)
        r_ = rates[i]
        rows_in = {q // r_ for q in need} | {q // r_ + (-1 if right else 1) for q in need}
        tconv[i] = min(q // r_ for q in need)       # transposed-conv op row m produces output rows m*r .. m*r + r - 1
        need = {m for m in rows_in if 0 <= m < lvl[i]}
    dec0 = min(need)
    return dec0, tconv, units, fin


@pytest.mark.parametrize("trim", ["both", "right"])
@pytest.mark.parametrize("rows0,rates", [(132, (8, 5, 4, 3)), (32, (8, 5, 4, 3)), (36, (2, 2)), (64, (3,)), (9, (4, 3))])
def test_tail_starts_match_dependency_closure(rows0, rates, trim):
    lvl = [rows0]
    for r in rates:
        lvl.append(lvl[-1] * r if trim == "right" else (lvl[-1] - 1) * r)
    n = lvl[-1]
    for skip in sorted({0, 1, 7, n // 3, n // 2, (3 * n) // 4, n - 1}):
        got = vocoder_tail_starts(rows0, rates, skip, trim)
        ref = brute_force(rows0, rates, skip, trim)
        assert got[3] == ref[3] == skip
        for i in range(len(rates)):
            for j in range(3):
                assert got[2][i][j] == ref[2][i][j], (skip, i, j)          # tight: exactly the first needed row
            assert got[1][i] <= ref[1][i] and ref[1][i] - got[1][i] <= 1, (skip, i)
        assert got[0] <= ref[0] and ref[0] - got[0] <= 1


def test_tail_starts_full_decode_is_zero():
    dec0, tconv, units, fin = vocoder_tail_starts(132, (8, 5, 4, 3), 0)
    assert dec0 == 0 and fin == 0 and all(t == 0 for t in tconv) and all(u == 0 for row in units for u in row)


def test_tail_starts_save_most_of_a_streaming_window():
    """33-frame window, 25 context frames: the vocoder only has to produce about a third of its rows."""
    rows0, rates = 132, (8, 5, 4, 3)
    n = rows0
    for r in rates:
        n = (n - 1) * r
    skip = int(round(25 * n / 33))
    dec0, tconv, units, fin = vocoder_tail_starts(rows0, rates, skip)
    assert fin == skip and dec0 >= 80          # of 132 rows (row 100 = first kept frame; the reach is ~3.5 frames)
    assert units[-1][0] > 0.7 * n              # last block (the most expensive one) starts near the kept samples


def test_speech_tokenizer_accepts_the_dict_batch_and_list_forms():
    """`speech_tokenizer.decode` is called with {"audio_codes": [B, T, Q]} by the reference's model (model.py:642) and with a list of
    {"audio_codes": [T, Q]} by its examples (examples/generate_with_embedding.py:98): both give one waveform per item."""
    import types

    import torch

    from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer

    class Dec:
        cfg = types.SimpleNamespace(sample_rate=24000)

        def decode(self, codes, skip_samples=0):
            assert codes.dim() == 2
            return codes[:, 0].float().repeat_interleave(4)[skip_samples:]

    tok = SpeechTokenizer(Dec())
    a, b = torch.arange(6).reshape(3, 2), torch.arange(10, 14).reshape(2, 2)
    wavs, sr = tok.decode({"audio_codes": a.unsqueeze(0)})
    assert sr == 24000 and len(wavs) == 1 and wavs[0].shape == (12,)
    wavs, sr = tok.decode([{"audio_codes": a}, {"audio_codes": b}])
    assert sr == 24000 and [w.shape[0] for w in wavs] == [12, 8] and wavs[1][0] == 10
    wavs, _ = tok.decode({"audio_codes": a, "skip_samples": 4})
    assert wavs[0].shape == (8,)
    assert tok.decode(a)[0][0].shape == (12,)
