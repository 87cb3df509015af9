"""Device timing of batched lock-step decode (n_streams = 1 .. 64) on synthetic weights."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("FQ3_VARIANT"): sys.path.insert(0, os.path.join(ROOT, "variants", os.environ["FQ3_VARIANT"]))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
name = sys.argv[1] if len(sys.argv) > 1 else "0.6B-Base"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = make_cfg(name)
w = make_weights(cfg, seed=0, norm_jitter=0.0)
counts = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "1,2,4,8,16,32,64".split(","))]
eng = make_engine(cfg, w, max_seq_len=int(os.environ.get("MAX_SEQ", "1024")), max_streams=max(counts), max_frames=2048)
print("lockstep group", eng.lockstep_group)
pol = SamplingPolicy(do_sample=True, temperature=0.9, top_k=50, repetition_penalty=1.05, min_new_tokens=10000)
sub = SubPolicy(do_sample=True, top_k=50, temperature=0.9)
ev = lambda: torch.cuda.Event(enable_timing=True)
for ns in counts:
    for s in range(ns):
        tie, tam, tth, tpe = synth_prompt(cfg, T=14, seed=1 + s)
        eng.set_text_conditioning(s, tth[0].cuda(), tpe.cuda())
        eng.prefill(s, tie[0].cuda(), 0, pol)
    eng.decode_frames(ns, 4, pol, sub)
    ts = []
    for _ in range(3):
        a, b = ev(), ev(); torch.cuda.synchronize(); a.record(); eng.decode_frames(ns, frames, pol, sub); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[1]
    st = [eng.status(s) for s in range(ns)]
    print(json.dumps(dict(model=name, n_streams=ns, ms_per_frame_step=ms / frames, aggregate_rtf=ns * 0.08 / (ms / frames / 1000), errors=[x.error for x in st],
                          frames=[x.n_frames for x in st])))
