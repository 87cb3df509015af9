"""BASELINE configs[2]: Qwen3-TTS-12Hz-1.7B-CustomVoice, speaker aiden, streaming chunk_size 8, bs 1 — TTFA and RTF through the
public API (random-init weights of the named architecture, 256 frames)."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
name = sys.argv[1] if len(sys.argv) > 1 else "synthetic://1.7B-CustomVoice"
m = FasterQwen3TTS.from_pretrained(name, device="cuda:0", dtype=torch.bfloat16, attn_implementation="eager", max_seq_len=2048, seed=0)
kw = dict(text="Hello world! This is a streaming test of the custom voice model.", speaker="aiden", language="English", chunk_size=8,
          max_new_tokens=256, min_new_tokens=256)
for _ in range(2):
    for _a in m.generate_custom_voice_streaming(**kw): pass
ttfa = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = m.generate_custom_voice_streaming(**kw); next(g); torch.cuda.synchronize()
    ttfa.append((time.perf_counter() - t0) * 1e3); g.close()
rtf = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
    for audio, sr, _t in m.generate_custom_voice_streaming(**kw): n += len(audio)
    torch.cuda.synchronize(); rtf.append(n / sr / (time.perf_counter() - t0))
print(json.dumps({"config": name + " speaker=aiden streaming chunk_size=8 bs=1, 256 frames, random-init weights",
                  "ttfa_ms": {"mean": float(np.mean(ttfa)), "std": float(np.std(ttfa))}, "rtf_e2e": float(np.median(rtf))}))
