"""BASELINE configs[3]: Qwen3-TTS-12Hz-1.7B-VoiceDesign, non-streaming, long text (~60 s of audio = 750 frames), through the public
API at bs = 1, plus the same prompt as 4 and as 16 requests through fast_generate_batch.  At 1.7B dims a lock-step group holds four
streams (a 6144-column activation row is 12 KB: the 16-row staging of the wide frame program does not fit next to the weight ring), so 16
requests run as four groups of four, one after the other.  Random-init weights of the named architecture."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
from qwen3_tts_cuda_graphs_b200.generate import fast_generate_batch
name = "synthetic://1.7B-VoiceDesign"
text = " ".join(["This is a long paragraph read by a designed voice, sentence number %d." % i for i in range(1, 13)])
instruct = "A calm, low-pitched male narrator with a slow pace."
frames = 750
m = FasterQwen3TTS.from_pretrained(name, device="cuda:0", dtype=torch.bfloat16, attn_implementation="eager", max_seq_len=2048, seed=0, max_streams=4)
kw = dict(text=text, instruct=instruct, language="English", max_new_tokens=frames, min_new_tokens=frames)
m.generate_voice_design(**kw)
ts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    audio, sr = m.generate_voice_design(**kw)
    torch.cuda.synchronize(); ts.append(len(audio[0]) / sr / (time.perf_counter() - t0))
out = {"config": name + f" non-streaming, {frames} frames (60 s), random-init weights", "rtf_bs1_e2e": float(np.median(ts))}
mm, talker, config, tie, tam, tth, tpe = m._prepare_generation_custom(text, "English", None, instruct)
reqs = [(tie, tam, tth, tpe)] * 4
def batched():
    codes, _ = fast_generate_batch(m.talker_graph, m.predictor_graph, reqs, max_new_tokens=frames, min_new_tokens=frames)
    n = 0
    for c in codes:
        a, sr_ = m._decode_full(mm, c)
        n += len(a[0])
    return n / sr_
batched()
torch.cuda.synchronize(); t0 = time.perf_counter()
a_s = batched() + batched()
torch.cuda.synchronize()
out["audio_s_per_s_4_streams"] = a_s / (time.perf_counter() - t0)
reqs = [(tie, tam, tth, tpe)] * 16
torch.cuda.synchronize(); t0 = time.perf_counter()
a_s = batched()
torch.cuda.synchronize()
out["audio_s_per_s_16_requests_in_groups_of_4"] = a_s / (time.perf_counter() - t0)
out["lockstep_group"] = m.model.engine.lockstep_group
print(json.dumps(out))
