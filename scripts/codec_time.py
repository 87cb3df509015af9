"""Device time of the codec decoder for the streaming window sizes (8 and 33 frames), 0.6B codec dims, synthetic weights."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from qwen3_tts_cuda_graphs_b200.config import preset
from qwen3_tts_cuda_graphs_b200.codec import SpeechTokenizer
cfg = preset("0.6B-Base")
dec = SpeechTokenizer.synthetic(cfg.codec, torch.device("cuda"), seed=1).decoder
g = torch.Generator().manual_seed(0)
for T in (8, 33):
    codes = torch.randint(0, cfg.codec.codebook_size, (T, cfg.codec.num_quantizers), generator=g).cuda()
    for _ in range(3): w = dec.decode(codes)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): w = dec.decode(codes)
    b.record(); torch.cuda.synchronize()
    print(f"T={T}: {a.elapsed_time(b)/10:.3f} ms per decode, sum {float(w.double().sum()):.6f}")
