"""Does the codec decode run beside a decode kernel that leaves SMs free?  (FQ3_GRID=128: 20 free SMs.)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer
cfg = make_cfg("0.6B-Base")
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=2048)
dec = SpeechTokenizer.synthetic(preset("0.6B-Base").codec, torch.device("cuda"), seed=1).decoder
pol = SamplingPolicy(do_sample=True, temperature=0.9, top_k=50, repetition_penalty=1.05, min_new_tokens=10000)
sub = SubPolicy(do_sample=True, top_k=50, temperature=0.9)
tie, tam, tth, tpe = synth_prompt(cfg, T=39)
eng.set_text_conditioning(0, tth[0].cuda(), tpe.cuda())
eng.prefill(0, tie[0].cuda(), 0, pol)
codes = torch.randint(0, 2048, (33, 16), generator=torch.Generator().manual_seed(0)).cuda()
skip = int(round(25 * dec.n_samples(33) / 33))
side = torch.cuda.Stream()
for _ in range(3):
    eng.decode_frames(1, 8, pol, sub); dec.decode(codes, skip)
    with torch.cuda.stream(side): dec.decode(codes, skip)
torch.cuda.synchronize()
def timed(fn, n=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def seq():
    eng.decode_frames(1, 8, pol, sub); dec.decode(codes, skip)
def ovl():
    with torch.cuda.stream(side): dec.decode(codes, skip)   # codec of the previous chunk on the side stream ...
    eng.decode_frames(1, 8, pol, sub)                        # ... while the next chunk is generated
def lm():
    eng.decode_frames(1, 8, pol, sub)
print(f"grid {eng.num_sms}: LM chunk alone {timed(lm):.3f} ms | LM then codec {timed(seq):.3f} ms | codec on a side stream + LM {timed(ovl):.3f} ms")
