"""Where the time to first audio goes (warm): prompt build, prefill, first chunk of frames, codec, host copy."""
import os, sys, time, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
from qwen3_tts_cuda_graphs_b200.streaming import fast_generate_streaming
model = FasterQwen3TTS.from_pretrained("synthetic://0.6B-Base", device="cuda:0", dtype=torch.bfloat16, attn_implementation="eager",
                                       max_seq_len=2048, seed=0)
ref_wav = bench.make_ref_wav()
kw = dict(text=bench.TEXT, language="English", ref_audio=ref_wav, ref_text=bench.REF_TEXT, chunk_size=8, max_new_tokens=64, min_new_tokens=64)
for _ in range(3):
    g = model.generate_voice_clone_streaming(**kw); next(g); g.close()
sync = torch.cuda.synchronize
def T(): sync(); return time.perf_counter()
rows = []
for _ in range(5):
    t0 = T()
    m, talker, tconf, tie, tam, tth, tpe, ref_codes = model._prepare_generation(bench.TEXT, ref_wav, bench.REF_TEXT, language="English", non_streaming_mode=True)
    t1 = T()
    stream = fast_generate_streaming(talker=talker, talker_input_embeds=tie, attention_mask=tam, trailing_text_hiddens=tth, tts_pad_embed=tpe,
                                     config=tconf, predictor_graph=model.predictor_graph, talker_graph=model.talker_graph, chunk_size=8,
                                     max_new_tokens=64, min_new_tokens=64)
    codes, timing = next(stream)
    t2 = T()
    gen = model._stream_audio(m, iter([(codes, timing)]), None, 8)
    audio, sr, _ = next(gen)
    t3 = T()
    rows.append(((t1 - t0) * 1e3, timing.get("prefill_ms", 0.0), (t2 - t1) * 1e3 - timing.get("prefill_ms", 0.0), (t3 - t2) * 1e3, (t3 - t0) * 1e3))
    stream.close()
print("prompt build | prefill | first 8 frames (+host) | codec + D2H | total   (ms)")
for r in rows: print("  " + "  ".join(f"{x:7.2f}" for x in r))
