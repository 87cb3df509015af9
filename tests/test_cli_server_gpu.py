"""The command line and the HTTP server end to end on the GPU (reference: faster_qwen3_tts/cli.py,
examples/openai_server.py) with the tiny synthetic preset: `clone` writes a playable WAV of the expected length, `serve --concurrency`
feeds stdin lines to the scheduler, POST /v1/audio/speech streams the same samples the scheduler hands out."""
import io
import struct
import sys
import wave

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def ref_wav(tmp_path_factory):
    p = tmp_path_factory.mktemp("audio") / "ref.wav"
    sr = 24000
    t = np.arange(int(1.0 * sr)) / sr
    pcm = (0.3 * np.sin(2 * np.pi * 180 * t) * 32767).astype(np.int16)
    with wave.open(str(p), "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes(pcm.tobytes())
    return str(p)


def _wav_info(path):
    with wave.open(path) as wf:
        return wf.getframerate(), wf.getnchannels(), wf.getsampwidth(), wf.getnframes()


def test_cli_clone_streaming_and_not(tmp_path, ref_wav, capsys):
    from qwen3_tts_cuda_graphs_b200 import cli

    for extra, name in (([], "a.wav"), (["--streaming", "--chunk-size", "8"], "b.wav")):
        out = str(tmp_path / name)
        cli.main(["clone", "--model", "tiny-Base", "--text", "Hello world, the quick brown fox.", "--language", "English", "--output", out,
                  "--ref-audio", ref_wav, "--ref-text", "reference clip", "--xvec-only", "--greedy", "--max-new-tokens", "20"] + extra)
        assert _wav_info(out) == (24000, 1, 2, 20 * 1920)
    assert "RTF" in capsys.readouterr().out


def test_cli_custom_lists_speakers_and_serve_uses_the_scheduler(tmp_path, monkeypatch, capsys):
    from qwen3_tts_cuda_graphs_b200 import cli

    cli.main(["custom", "--model", "tiny-CustomVoice", "--text", "x", "--output", str(tmp_path / "x.wav"), "--list-speakers"])
    assert "aiden" in capsys.readouterr().out.split()
    monkeypatch.setattr(sys, "stdin", io.StringIO("First line to speak.\n\nSecond line, a little longer than the first.\nquit\nnever read\n"))
    cli.main(["serve", "--mode", "custom", "--model", "tiny-CustomVoice", "--speaker", "aiden", "--language", "English", "--greedy",
              "--max-new-tokens", "18", "--output-dir", str(tmp_path / "out"), "--concurrency", "3"])
    assert _wav_info(str(tmp_path / "out" / "out_0001.wav"))[3] == 18 * 1920
    assert _wav_info(str(tmp_path / "out" / "out_0002.wav"))[3] == 18 * 1920
    assert not (tmp_path / "out" / "out_0003.wav").exists()


def test_http_speech_streams_what_the_scheduler_decodes(ref_wav):
    from starlette.testclient import TestClient

    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS, server
    from qwen3_tts_cuda_graphs_b200.serving import BatchScheduler

    tts = FasterQwen3TTS.from_pretrained("tiny-Base", device="cuda", dtype=torch.bfloat16, max_seq_len=256, max_streams=4)
    tts.predictor_graph.do_sample = False
    voices = {"alloy": {"ref_audio": ref_wav, "ref_text": "", "language": "English", "max_new_tokens": 24, "do_sample": False}}
    try:
        with BatchScheduler(tts, chunk_frames=8) as sched:
            client = TestClient(server.create_app([sched], voices, "alloy", sample_rate=tts.sample_rate))
            r = client.post("/v1/audio/speech", json={"input": "Hello world!", "voice": "alloy", "response_format": "wav"})
            assert r.status_code == 200 and r.content[:4] == b"RIFF" and struct.unpack("<I", r.content[24:28])[0] == 24000
            pcm = np.frombuffer(r.content[44:], dtype="<i2")
            assert pcm.size == 24 * 1920
            r2 = client.post("/v1/audio/speech", json={"input": "Hello world!", "voice": "alloy", "response_format": "pcm"})
            assert np.array_equal(np.frombuffer(r2.content, dtype="<i2"), pcm)  # greedy: the same utterance twice
        # the bytes are the public streaming API's samples, converted the way the reference's server converts them
        want = np.concatenate([a for a, _, _ in tts.generate_voice_clone_streaming("Hello world!", "English", ref_wav, "", max_new_tokens=24,
                                                                                   do_sample=False, chunk_size=8, non_streaming_mode=False)])
        assert np.array_equal(pcm, np.clip(want * 32768, -32768, 32767).astype(np.int16))
    finally:
        tts.model.engine.close()
