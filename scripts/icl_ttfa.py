"""ICL voice clone (xvec_only=False: the reference clip's codes in the prompt, SURVEY.md §8d C2 "ICL variant", prompt T ~ 240) through the
public streaming API: time to first audio (chunk 8) and real-time factor of a 256-frame utterance, warm."""
import os, sys, time, wave
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
model = FasterQwen3TTS.from_pretrained("synthetic://0.6B-Base", device="cuda:0", dtype=torch.bfloat16, max_seq_len=2048, seed=0)
path = "/tmp/icl_ref.wav"
sr = 24000
t = np.arange(int(13.5 * sr)) / sr   # 13.5 s clip (+0.5 s silence appended by the API) = 175 reference frames
with wave.open(path, "wb") as wf:
    wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(sr)
    wf.writeframes((0.3 * np.sin(2 * np.pi * 220 * t) * 32767).astype(np.int16).tobytes())
ref_text = " ".join(["word"] * 48)
kw = dict(text=bench.TEXT, language="English", ref_audio=path, ref_text=ref_text, chunk_size=8, xvec_only=False, non_streaming_mode=True)
for mode in ("icl", "xvec"):
    k = dict(kw, xvec_only=(mode == "xvec"))
    for _ in range(3):
        g = model.generate_voice_clone_streaming(max_new_tokens=16, min_new_tokens=16, **k); next(g); g.close()
    ttfa = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g = model.generate_voice_clone_streaming(max_new_tokens=64, min_new_tokens=64, **k)
        next(g); ttfa.append((time.perf_counter() - t0) * 1e3); g.close()
    best = None
    for _ in range(3):  # the first pass builds the launch plans of the sizes only a full utterance meets
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = sum(len(a) for a, _, _ in model.generate_voice_clone_streaming(max_new_tokens=256, min_new_tokens=256, **k))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    dt = best
    m, talker, tconf, tie, *_ = model._prepare_generation(bench.TEXT, path, ref_text, language="English", xvec_only=(mode == "xvec"), non_streaming_mode=True)
    print(f"{mode}: prompt rows {tie.shape[1]}, TTFA {np.mean(ttfa):.2f} +- {np.std(ttfa):.2f} ms, 256 frames in {dt * 1e3:.1f} ms -> RTF {n / 24000 / dt:.1f}")
