"""HTTP layer of the serving row against a stand-in backend: the OpenAI endpoint's contract
(examples/openai_server.py:219-265 — streamed WAV with an unknown-length header, pcm, 400s for empty input / unknown voice /
unsupported format, default-voice fallback) and the demo's SSE protocol (demo/server.py:332-541: queued / chunk / done)."""
import base64
import json
import struct

import numpy as np
import pytest
from starlette.testclient import TestClient

from qwen3_tts_cuda_graphs_b200 import server
from qwen3_tts_cuda_graphs_b200.serving import RequestHandle, TTSRequest, _DONE


class StubBackend:
    """submit() -> a handle that already holds three 0.1 s chunks (a ramp, so order is checkable)."""

    def __init__(self, fail=False):
        self.requests = []
        self.fail = fail

    def submit(self, req: TTSRequest):
        h = RequestHandle(req, len(self.requests))
        self.requests.append(req)
        if self.fail:
            h.q.put(RuntimeError("boom"))
        else:
            for i in range(3):
                h.q.put((np.full(2400, 0.1 * (i + 1), dtype=np.float32), 24000, {"chunk_index": i, "is_final": i == 2}))
            h.finish_reason = "stop"
        h.q.put(_DONE)
        return h


VOICES = {"alloy": {"ref_audio": "alloy.wav", "ref_text": "hi", "language": "English"},
          "aiden": {"speaker": "aiden", "language": "English"}}


@pytest.fixture()
def client():
    b0, b1 = StubBackend(), StubBackend()
    app = server.create_app([b0, b1], VOICES, "alloy")
    return TestClient(app), b0, b1


def test_health_and_voices(client):
    c, *_ = client
    assert c.get("/health").json()["status"] == "ok"
    assert c.get("/v1/voices").json() == {"voices": ["aiden", "alloy"], "default": "alloy"}


def test_speech_streams_a_wav_with_unknown_length_header(client):
    c, b0, b1 = client
    r = c.post("/v1/audio/speech", json={"model": "tts-1", "input": "Hello!", "voice": "alloy", "response_format": "wav"})
    assert r.status_code == 200 and r.headers["content-type"] == "audio/wav"
    body = r.content
    assert body[:4] == b"RIFF" and body[8:16] == b"WAVEfmt " and struct.unpack("<I", body[40:44])[0] == 0xFFFFFFFF
    assert struct.unpack("<IHHIIHH", body[16:36]) == (16, 1, 1, 24000, 48000, 2, 16)
    pcm = np.frombuffer(body[44:], dtype="<i2")
    assert pcm.size == 7200 and abs(pcm[0] - 3276) <= 1 and abs(pcm[-1] - 9830) <= 1
    req = (b0.requests + b1.requests)[0]
    assert req.kind == "voice_clone" and req.ref_audio == "alloy.wav" and req.ref_text == "hi" and req.language == "English"
    assert req.non_streaming_mode is False  # examples/openai_server.py:190
    assert c.app.state.dispatcher.in_flight == [0, 0]  # released when the stream ended


def test_pcm_custom_voice_and_least_loaded_dispatch(client):
    c, b0, b1 = client
    r = c.post("/v1/audio/speech", json={"input": "Hi", "voice": "aiden", "response_format": "pcm"})
    assert r.status_code == 200 and r.headers["content-type"] == "audio/pcm" and len(r.content) == 7200 * 2
    c.post("/v1/audio/speech", json={"input": "Hi again", "voice": "aiden", "response_format": "pcm"})
    reqs = b0.requests + b1.requests
    assert len(reqs) == 2 and all(q.kind == "custom_voice" and q.speaker == "aiden" for q in reqs)
    d = server.Dispatcher([b0, b1])
    h0, h1 = d.submit(TTSRequest("a")), d.submit(TTSRequest("b"))
    assert {h0._backend_index, h1._backend_index} == {0, 1}  # the second goes to the idle replica


def test_errors_mirror_the_reference(client):
    c, *_ = client
    assert c.post("/v1/audio/speech", json={"input": "   ", "voice": "alloy"}).status_code == 400
    assert c.post("/v1/audio/speech", json={"input": "x", "voice": "alloy", "response_format": "flac"}).status_code == 400
    r = c.post("/v1/audio/speech", json={"input": "x", "voice": "alloy", "response_format": "mp3"})
    assert r.status_code == 400 and "requires pydub" in r.json()["detail"]  # examples/openai_server.py:126-129 (pydub is not in this image)
    # unknown voice falls back to the default one (examples/openai_server.py:150-157) ...
    assert c.post("/v1/audio/speech", json={"input": "x", "voice": "nobody"}).status_code == 200
    # ... and is a 400 when there is no default
    app = server.create_app([StubBackend()], VOICES, None)
    assert TestClient(app).post("/v1/audio/speech", json={"input": "x", "voice": "nobody"}).status_code == 400


def test_sse_protocol(client):
    c, *_ = client
    r = c.post("/generate/stream", data={"text": "Hello there", "mode": "custom", "speaker": "aiden", "language": "English"})
    assert r.status_code == 200 and r.headers["content-type"].startswith("text/event-stream")
    msgs = [json.loads(line[6:]) for line in r.text.splitlines() if line.startswith("data: ")]
    assert [m["type"] for m in msgs] == ["chunk", "chunk", "chunk", "done"]  # "queued" only when somebody is ahead (demo/server.py:513)
    wav = base64.b64decode(msgs[0]["audio_b64"])
    assert wav[:4] == b"RIFF" and struct.unpack("<I", wav[40:44])[0] == 4800 and msgs[0]["sample_rate"] == 24000
    assert msgs[-1]["total_audio_s"] == pytest.approx(0.3, abs=1e-3)
    r = c.post("/generate/stream", data={"text": "x", "mode": "nonsense"})
    assert r.status_code == 400


def test_sse_reports_generation_errors_in_band():
    app = server.create_app([StubBackend(fail=True)], VOICES, "alloy")
    r = TestClient(app).post("/generate/stream", data={"text": "x", "mode": "voice_clone"})
    msgs = [json.loads(line[6:]) for line in r.text.splitlines() if line.startswith("data: ")]
    assert msgs[-1] == {"type": "error", "message": "boom"}


def test_cli_parser_has_the_reference_subcommands_and_writes_wav(tmp_path):
    import wave

    from qwen3_tts_cuda_graphs_b200 import cli

    p = cli.build_parser()
    a = p.parse_args(["clone", "--model", "m", "--text", "t", "--output", "o.wav", "--ref-audio", "a.wav", "--ref-text", "r", "--greedy",
                      "--streaming", "--no-non-streaming-mode"])
    assert (a.greedy, a.streaming, a.non_streaming_mode, a.chunk_size, a.max_new_tokens) == (True, True, False, 8, 2048)
    stem, kw = cli._call(None, "clone", "t", a)
    assert stem == "generate_voice_clone" and kw["do_sample"] is False and kw["ref_audio"] == "a.wav" and kw["non_streaming_mode"] is False
    a = p.parse_args(["serve", "--mode", "design", "--model", "m", "--instruct", "calm", "--concurrency", "8"])
    assert a.concurrency == 8 and cli._call(None, "design", "t", a)[0] == "generate_voice_design"
    a = p.parse_args(["custom", "--model", "m", "--text", "t", "--output", "o.wav", "--list-speakers"])
    assert a.list_speakers
    out = str(tmp_path / "d" / "x.wav")
    cli.write_audio(out, np.linspace(-1, 1, 480, dtype=np.float32), 24000)
    with wave.open(out) as wf:
        assert (wf.getframerate(), wf.getnchannels(), wf.getsampwidth(), wf.getnframes()) == (24000, 1, 2, 480)


def test_warm_up_runs_every_voice_on_every_replica_and_survives_a_broken_voice():
    b0, b1 = StubBackend(), StubBackend(fail=True)
    dt = server.warm_up([b0, b1], VOICES, rounds=2, frames=12)
    assert dt >= 0
    for b in (b0, b1):
        assert len(b.requests) == 2 * len(VOICES)
        assert {r.kind for r in b.requests} == {"voice_clone", "custom_voice"} and all(r.max_new_tokens == 12 for r in b.requests)


def test_dead_replicas_are_routed_around_and_all_dead_is_a_503():
    b0, b1 = StubBackend(), StubBackend()
    b0.healthy = False  # serving.BatchScheduler.healthy after its loop died (device fault)
    c = TestClient(server.create_app([b0, b1], VOICES, "alloy"))
    assert c.get("/health").json()["status"] == "degraded" and c.get("/health").json()["healthy"] == [False, True]
    for _ in range(3):
        assert c.post("/v1/audio/speech", json={"input": "Hi", "voice": "alloy"}).status_code == 200
    assert len(b0.requests) == 0 and len(b1.requests) == 3
    b1.healthy = False
    assert c.get("/health").json()["status"] == "down"
    r = c.post("/v1/audio/speech", json={"input": "Hi", "voice": "alloy"})
    assert r.status_code == 503 and "no healthy backend" in r.json()["detail"]
    r = c.post("/generate/stream", data={"text": "Hi", "voice": "alloy"})
    assert r.status_code == 503
    assert c.app.state.dispatcher.in_flight == [0, 0]


def test_sse_says_queued_when_requests_are_ahead(client):
    c, *_ = client
    c.app.state.dispatcher.in_flight[0] = 3  # three utterances already running on replica 0
    r = c.post("/generate/stream", data={"text": "Hi", "mode": "voice_design", "instruct": "a calm voice"})
    msgs = [json.loads(line[6:]) for line in r.text.splitlines() if line.startswith("data: ")]
    assert msgs[0] == {"type": "queued", "position": 3} and msgs[-1]["type"] == "done"


def test_uploaded_reference_audio_becomes_a_content_addressed_file(client, tmp_path):
    """demo/server.py:366-373: an uploaded clip is written once under its content hash, so a repeated upload is the same
    `ref_audio` path (the model's voice-prompt cache key)."""
    import os

    c, b0, b1 = client
    clip = b"RIFF" + bytes(range(256)) * 4
    paths = []
    for _ in range(2):
        r = c.post("/generate/stream", data={"text": "Hello", "mode": "voice_clone", "ref_text": "what the clip says", "xvec_only": "false"},
                   files={"ref_audio": ("my voice.wav", clip, "audio/wav")})
        assert r.status_code == 200
    reqs = b0.requests + b1.requests
    assert len(reqs) == 2 and reqs[0].ref_audio == reqs[1].ref_audio and reqs[0].kind == "voice_clone"
    assert reqs[0].ref_text == "what the clip says" and reqs[0].xvec_only is False
    with open(reqs[0].ref_audio, "rb") as f:
        assert f.read() == clip
    os.unlink(reqs[0].ref_audio)
    # without an upload: the named preset / voice, else the default voice
    c.post("/generate/stream", data={"text": "Hello", "ref_preset": "alloy"})
    assert (b0.requests + b1.requests)[-1].ref_audio == "alloy.wav" or b1.requests[-1].ref_audio == "alloy.wav"


def test_demo_limits_and_one_shot_endpoint(client, monkeypatch):
    c, b0, b1 = client
    r = c.post("/generate/stream", data={"text": "x" * (server.MAX_TEXT_CHARS + 1)})
    assert r.status_code == 400 and "Text too long" in r.json()["detail"]
    monkeypatch.setattr(server, "MAX_AUDIO_BYTES", 64)
    r = c.post("/generate", data={"text": "Hi"}, files={"ref_audio": ("big.wav", b"0" * 65, "audio/wav")})
    assert r.status_code == 400 and "Audio file too large" in r.json()["detail"]
    r = c.post("/generate", data={"text": "Hi", "mode": "custom", "speaker": "aiden", "temperature": "0.5", "top_k": "20"})
    assert r.status_code == 200
    body = r.json()
    wav = base64.b64decode(body["audio_b64"])
    assert wav[:4] == b"RIFF" and struct.unpack("<I", wav[40:44])[0] == 7200 * 2 and body["sample_rate"] == 24000
    assert body["metrics"]["audio_duration_s"] == pytest.approx(0.3, abs=1e-3) and set(body["metrics"]) == {"total_ms", "audio_duration_s", "rtf"}
    req = (b0.requests + b1.requests)[-1]
    assert req.kind == "custom_voice" and req.temperature == 0.5 and req.top_k == 20
    assert c.app.state.dispatcher.in_flight == [0, 0]
    failing = TestClient(server.create_app([StubBackend(fail=True)], VOICES, "alloy"))
    assert failing.post("/generate", data={"text": "Hi"}).status_code == 500


def test_status_lists_preset_voices(client):
    c, *_ = client
    st = c.get("/status").json()
    assert st["loaded"] is True and st["queue_depth"] == 0 and st["healthy"] == [True, True]
    assert st["preset_refs"] == [{"id": "alloy", "label": "alloy", "ref_text": "hi"}]


def test_replica_placement():
    assert server.replica_devices(1) == ["cuda:0"] and server.replica_devices(1, "cuda:3") == ["cuda:3"]
    assert server.replica_devices(4, "cuda:3") == ["cuda:0", "cuda:1", "cuda:2", "cuda:3"]


def test_mp3_is_one_encode_of_the_whole_utterance_when_pydub_exists(client, monkeypatch):
    """examples/openai_server.py:242-259 with a stand-in for pydub.AudioSegment (pydub + ffmpeg are not in this image)."""
    import sys
    import types

    seen = {}

    class Segment:
        def __init__(self, raw, frame_rate, sample_width, channels):
            seen.update(n=len(raw), frame_rate=frame_rate, sample_width=sample_width, channels=channels)

        def export(self, buf, format):
            buf.write(b"ID3" + format.encode())

    monkeypatch.setitem(sys.modules, "pydub", types.SimpleNamespace(AudioSegment=Segment))
    c, b0, b1 = client
    r = c.post("/v1/audio/speech", json={"input": "Hello", "voice": "alloy", "response_format": "mp3"})
    assert r.status_code == 200 and r.headers["content-type"] == "audio/mpeg" and r.content == b"ID3mp3"
    assert seen == dict(n=7200 * 2, frame_rate=24000, sample_width=2, channels=1)
    assert c.app.state.dispatcher.in_flight == [0, 0]


def test_demo_side_endpoints(tmp_path):
    clip = tmp_path / "alloy.wav"
    clip.write_bytes(b"RIFFxxxxWAVE")
    voices = {"alloy": {"ref_audio": str(clip), "ref_text": "hi"}, "aiden": {"speaker": "aiden"}}
    c = TestClient(server.create_app([StubBackend()], voices, "alloy", model_name="synthetic://0.6B-Base"))
    r = c.get("/preset_ref/alloy").json()
    assert r["filename"] == "alloy.wav" and r["ref_text"] == "hi" and base64.b64decode(r["audio_b64"]) == b"RIFFxxxxWAVE"
    assert c.get("/preset_ref/aiden").status_code == 404 and c.get("/preset_ref/nobody").status_code == 404
    assert c.post("/load", data={"model_id": "synthetic://0.6B-Base"}).json()["model"] == "synthetic://0.6B-Base"
    r = c.post("/load", data={"model_id": "Qwen/Qwen3-TTS-12Hz-1.7B-Base"})
    assert r.status_code == 400 and "--model" in r.json()["detail"]
    assert c.post("/transcribe", files={"audio": ("a.wav", b"x", "audio/wav")}).status_code == 503
    assert c.get("/status").json()["model"] == "synthetic://0.6B-Base"


def test_a_response_that_never_streams_still_gives_the_slot_back():
    """A client that goes away before the first byte never starts the body generator; the response's background task settles
    the request (cancel + in-flight count) all the same."""
    import asyncio

    class NeverEnding:
        def __init__(self):
            self.handles = []

        def submit(self, req):
            h = RequestHandle(req, len(self.handles))  # nothing is ever put on its queue
            self.handles.append(h)
            return h

    b = NeverEnding()
    app = server.create_app([b], VOICES, "alloy")
    route = next(r for r in app.routes if getattr(r, "path", "") == "/v1/audio/speech")
    resp = asyncio.run(route.endpoint(server.SpeechRequest(input="Hello", voice="alloy")))
    assert app.state.dispatcher.in_flight == [1]
    asyncio.run(resp.background())  # what Starlette runs when the response ends, streamed or not
    assert app.state.dispatcher.in_flight == [0] and b.handles[0].cancelled
    asyncio.run(resp.background())  # idempotent
    assert app.state.dispatcher.in_flight == [0]


def test_http_front_on_the_real_scheduler_with_engine_doubles():
    """server.create_app over serving.BatchScheduler itself (engine / model doubles of test_serving_cpu.py): two concurrent HTTP
    requests share launches, the WAV body holds every frame once and in order, the slots and in-flight counts come back."""
    import threading

    from test_serving_cpu import SPF, FakeEngine, FakeTTS

    from qwen3_tts_cuda_graphs_b200.serving import BatchScheduler

    eng = FakeEngine(max_streams=4, frame_sleep=0.001)
    voices = {"alloy": {"ref_audio": "v.wav", "language": "English", "max_new_tokens": 240, "do_sample": False}}
    with BatchScheduler(FakeTTS(eng), chunk_frames=8) as sched:
        c = TestClient(server.create_app([sched], voices, "alloy"))
        out = {}

        def post(key):
            out[key] = c.post("/v1/audio/speech", json={"input": f"{key},1000000000", "voice": "alloy", "response_format": "pcm"})

        ts = [threading.Thread(target=post, args=(k,)) for k in (5, 6)]
        [t.start() for t in ts]
        [t.join(60) for t in ts]
        for k in (5, 6):
            assert out[k].status_code == 200
            pcm = np.frombuffer(out[k].content, dtype="<i2")
            assert pcm.size == 240 * SPF  # 30 launches of 8 ms each: the two requests overlap whatever the threads' start skew
            # FakeDecoder: frame i -> 1920 samples of value i (clipped to int16 full scale by the PCM conversion)
            assert np.array_equal(pcm[::SPF], np.clip(np.arange(240) * 32768.0, -32768, 32767).astype(np.int16))
        assert c.get("/health").json() == {"status": "ok", "model_loaded": True, "backends": 1, "healthy": [True], "in_flight": [0]}
        r = c.post("/generate", data={"text": "9,5", "voice": "alloy"})  # EOS after five frames
        assert r.status_code == 200 and r.json()["metrics"]["audio_duration_s"] == pytest.approx(5 * 0.08)
    assert any(len(live) == 2 for _, live, _ in eng.launches)
