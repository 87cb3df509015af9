"""Frame cost of the decode kernel by policy (greedy / sampling) and context length (attention splits), 0.6B dims."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
name = sys.argv[1] if len(sys.argv) > 1 else "0.6B-Base"
cfg = make_cfg(name)
w = make_weights(cfg, seed=0, norm_jitter=0.0)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=2048)
ev = lambda: torch.cuda.Event(enable_timing=True)
pols = {
    "greedy": (SamplingPolicy(do_sample=False, repetition_penalty=1.0, min_new_tokens=10000), SubPolicy(do_sample=False)),
    "sample": (SamplingPolicy(do_sample=True, temperature=0.9, top_k=50, repetition_penalty=1.05, min_new_tokens=10000),
               SubPolicy(do_sample=True, top_k=50, temperature=0.9)),
}
for T in [int(x) for x in os.environ.get("FC_T", "14,39,240").split(",")]:
    tie, tam, tth, tpe = synth_prompt(cfg, T=T)
    eng.set_text_conditioning(0, tth[0].cuda(), tpe.cuda())
    tiec = tie[0].cuda()
    for pname, (pol, sub) in pols.items():
        eng.prefill(0, tiec, 0, pol)
        eng.decode_frames(1, 8, pol, sub)
        out = []
        for seg in range(4):  # 4 x 64 frames: the context grows by 64 per segment
            ts = []
            a, b = ev(), ev(); torch.cuda.synchronize(); a.record()
            for _ in range(8): eng.decode_frames(1, 8, pol, sub)
            b.record(); torch.cuda.synchronize()
            out.append(a.elapsed_time(b) / 64 * 1000)
        st = eng.status(0)
        print(f"{name} T={T:3d} {pname:6s}: us/frame over contexts +8..+72 | +72..+136 | +136..+200 | +200..+264: " + "  ".join(f"{x:7.1f}" for x in out) + f"   (frames {st.n_frames}, err {st.error})", flush=True)
