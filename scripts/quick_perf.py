"""Quick device timing of the decode engine on synthetic weights (CUDA events), for development."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("FQ3_VARIANT"): sys.path.insert(0, os.path.join(ROOT, "variants", os.environ["FQ3_VARIANT"]))  # A/B builds (scripts/build_variant.sh)
from helpers import make_cfg, make_weights, make_engine, synth_prompt
from qwen3_tts_cuda_graphs_b200.engine import SamplingPolicy, SubPolicy
from qwen3_tts_cuda_graphs_b200.weights import param_bytes

name = sys.argv[1] if len(sys.argv) > 1 else "0.6B-Base"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 14
cfg = make_cfg(name)
t0 = time.time(); w = make_weights(cfg, seed=0, norm_jitter=0.0); print("weights", time.time() - t0, flush=True)
eng = make_engine(cfg, w, max_seq_len=2048, max_frames=2048)
pb = param_bytes(cfg)
tie, tam, tth, tpe = synth_prompt(cfg, T=T)
pol = SamplingPolicy(do_sample=False, repetition_penalty=1.0); sub = SubPolicy(do_sample=False)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, n=5):
    ts = []
    for _ in range(n):
        a, b = ev(), ev(); torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]
eng.set_text_conditioning(0, tth[0].cuda(), tpe.cuda())
tiec = tie[0].cuda()
pf = timed(lambda: eng.prefill(0, tiec, 0, pol))
print(f"prefill T={T}: best {pf[0]:.3f} ms median {pf[1]:.3f} ms")
x = torch.randn(cfg.talker.hidden_size).to(torch.bfloat16).cuda()
ts = timed(lambda: eng.talker_step(0, x, T, want_logits=False), 10)
print(f"talker step: best {ts[0]*1000:.1f} us median {ts[1]*1000:.1f} us -> {pb['talker_step']/ts[1]/1e6:.0f} GB/s")
pin = torch.randn(2, cfg.talker.hidden_size).to(torch.bfloat16).cuda()
ps = timed(lambda: eng.predictor_run(0, pin, sub), 10)
pbytes = 15 * pb['predictor_pass'] + pb['predictor_heads']
print(f"predictor: best {ps[0]*1000:.1f} us median {ps[1]*1000:.1f} us -> {pbytes/ps[1]/1e6:.0f} GB/s (streaming bytes)")
def run():
    eng.prefill(0, tiec, 0, pol); eng.decode_frames(1, frames, pol, sub)
fr = timed(run, 5)
st = eng.status(0)
per = (fr[1] - pf[1]) / max(st.n_frames, 1)
print(f"frames={st.n_frames} err={st.error}: total median {fr[1]:.2f} ms, {per*1000:.1f} us/frame -> RTF {0.08/(per/1000):.1f}, "
      f"{pb['frame_streaming']/per/1e6:.0f} GB/s of streaming bytes ({pb['frame_streaming']/1e6:.1f} MB/frame)")
print(json.dumps(dict(model=name, prefill_ms=pf[1], talker_us=ts[1]*1000, predictor_us=ps[1]*1000, frame_us=per*1000, sms=eng.num_sms)))
