"""Command line (reference `faster_qwen3_tts/cli.py`) on CPU: the flag table equals the reference's
(tests/golden/cli_flags.json, written from the reference's own build_parser by tests/golden/make_cli_golden.py), every mode calls
the API method the reference calls with the same keyword arguments (cli.py:36-183), and the stdin `serve` loop writes one WAV per
line — one after the other, or through the continuous-batching scheduler with --concurrency."""
import io
import json
import os
import sys
import wave

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_cli_golden import flag_table  # noqa: E402

from qwen3_tts_cuda_graphs_b200 import cli  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cli_flags.json")


def test_flags_defaults_and_required_match_the_reference():
    with open(GOLDEN) as f:
        ref = json.load(f)
    ours = flag_table(cli.build_parser())
    missing = sorted(set(ref) - set(ours))
    assert not missing, f"reference flags missing here: {missing}"
    for k, v in ref.items():
        assert ours[k] == v, (k, ours[k], v)
    extra = sorted(set(ours) - set(ref))
    assert extra == ["serve --concurrency"], extra  # the one addition (lock-step utterances)


class FakeModel:
    """Records the API calls; audio = 0.25 s of a constant whose value encodes the call count."""

    def __init__(self):
        self.calls = []
        self.sample_rate = 24000

    def _audio(self):
        return np.full(6000, 0.01 * len(self.calls), dtype=np.float32)

    def __getattr__(self, name):
        if not name.startswith("generate_"):
            raise AttributeError(name)

        def call(**kw):
            self.calls.append((name, kw))
            if name.endswith("_streaming"):
                a = self._audio()
                return iter([(a[:2000], 24000, {}), (a[2000:], 24000, {})])
            return [self._audio()], 24000
        return call


def _run(monkeypatch, argv, stdin=None):
    model = FakeModel()
    loaded = []

    def fake_load(model_id, device, dtype, max_streams=1):
        loaded.append((model_id, device, dtype, max_streams))
        return model
    monkeypatch.setattr(cli, "_load_model", fake_load)
    if stdin is not None:
        monkeypatch.setattr(sys, "stdin", io.StringIO(stdin))
    cli.main(argv)
    return model, loaded


def _read_wav(path):
    with wave.open(str(path), "rb") as wf:
        assert wf.getnchannels() == 1 and wf.getsampwidth() == 2
        return np.frombuffer(wf.readframes(wf.getnframes()), dtype="<i2"), wf.getframerate()


def test_clone_one_shot_and_streaming(monkeypatch, tmp_path, capsys):
    out = tmp_path / "sub" / "a.wav"
    base = ["clone", "--model", "M", "--text", "Hello.", "--language", "English", "--output", str(out), "--ref-audio", "r.wav",
            "--ref-text", "ref"]
    model, loaded = _run(monkeypatch, base + ["--xvec-only", "--greedy", "--top-k", "20"])
    assert loaded == [("M", "cuda", "bf16", 1)]
    name, kw = model.calls[0]
    assert name == "generate_voice_clone"
    assert kw == dict(text="Hello.", language="English", max_new_tokens=2048, temperature=0.9, top_k=20, do_sample=False,
                      repetition_penalty=1.05, ref_audio="r.wav", ref_text="ref", xvec_only=True, non_streaming_mode=True)
    pcm, sr = _read_wav(out)
    assert sr == 24000 and pcm.size == 6000 and abs(int(pcm[0]) - 328) <= 1
    assert "Wrote" in capsys.readouterr().out
    model, _ = _run(monkeypatch, base + ["--streaming", "--chunk-size", "4", "--no-non-streaming-mode"])
    name, kw = model.calls[0]
    assert name == "generate_voice_clone_streaming" and kw["chunk_size"] == 4 and kw["non_streaming_mode"] is False
    assert kw["xvec_only"] is False and kw["do_sample"] is True
    assert _read_wav(out)[0].size == 6000  # the chunks were concatenated


def test_custom_and_design(monkeypatch, tmp_path):
    out = tmp_path / "b.wav"
    model, _ = _run(monkeypatch, ["custom", "--model", "M", "--text", "Hi", "--output", str(out), "--speaker", "aiden", "--instruct", "calm"])
    name, kw = model.calls[0]
    assert name == "generate_custom_voice" and kw["speaker"] == "aiden" and kw["instruct"] == "calm" and kw["language"] == "Auto"
    with pytest.raises(SystemExit) as e:  # cli.py: --speaker is required unless --list-speakers
        _run(monkeypatch, ["custom", "--model", "M", "--text", "Hi", "--output", str(out)])
    assert e.value.code == 2
    model, _ = _run(monkeypatch, ["design", "--model", "M", "--text", "Hi", "--output", str(out), "--instruct", "a deep voice", "--streaming"])
    name, kw = model.calls[0]
    assert name == "generate_voice_design_streaming" and kw["instruct"] == "a deep voice" and kw["chunk_size"] == 8


def test_serve_reads_stdin_lines_until_quit(monkeypatch, tmp_path):
    outdir = tmp_path / "outs"
    model, loaded = _run(monkeypatch, ["serve", "--mode", "clone", "--model", "M", "--ref-audio", "r.wav", "--ref-text", "ref",
                                       "--output-dir", str(outdir)], stdin="first line\n\nsecond line\nquit\nnever reached\n")
    assert [kw["text"] for _, kw in model.calls] == ["first line", "second line"]
    assert sorted(os.listdir(outdir)) == ["out_0001.wav", "out_0002.wav"]
    assert loaded[0][3] == 1
    for mode, flag in (("clone", "--ref-audio"), ("custom", "--speaker"), ("design", "--instruct")):
        with pytest.raises(SystemExit) as e:  # each mode's required argument (cli.py:191-199)
            _run(monkeypatch, ["serve", "--mode", mode, "--model", "M"], stdin="x\n")
        assert e.value.code == 2


def test_serve_concurrency_goes_through_the_scheduler(monkeypatch, tmp_path, capsys):
    """--concurrency N: lines become serving.TTSRequest objects of the mode's kind; a failing line is reported, the others are
    written."""
    from qwen3_tts_cuda_graphs_b200 import serving

    submitted = []

    class FakeHandle:
        def __init__(self, req):
            self.req = req

        def result(self):
            if "bad" in self.req.text:
                raise RuntimeError("Input is too long")
            return np.full(4800, 0.5, dtype=np.float32), 24000

    class FakeScheduler:
        def __init__(self, tts, chunk_frames, max_concurrent):
            self.args = (chunk_frames, max_concurrent)
            self.stopped = False
            FakeScheduler.last = self

        def start(self):
            return self

        def submit(self, req):
            submitted.append(req)
            return FakeHandle(req)

        def stop(self):
            self.stopped = True

    monkeypatch.setattr(serving, "BatchScheduler", FakeScheduler)
    outdir = tmp_path / "outs"
    _, loaded = _run(monkeypatch, ["serve", "--mode", "custom", "--model", "M", "--speaker", "aiden", "--language", "English",
                                   "--output-dir", str(outdir), "--concurrency", "4", "--chunk-size", "12", "--temperature", "0.7"],
                     stdin="one\nbad one\nthree\n")
    assert loaded[0][3] == 4 and FakeScheduler.last.args == (12, 4) and FakeScheduler.last.stopped
    assert [r.kind for r in submitted] == ["custom_voice"] * 3 and submitted[0].speaker == "aiden" and submitted[0].temperature == 0.7
    assert submitted[0].language == "English" and submitted[0].max_new_tokens == 2048
    assert sorted(os.listdir(outdir)) == ["out_0001.wav", "out_0003.wav"]
    assert "Input is too long" in capsys.readouterr().err
