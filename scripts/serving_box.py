"""The serving layer across the GPUs of one box, in ONE process: one model replica + serving.BatchScheduler per GPU (server.build_backends),
requests handed to the least-loaded replica (server.Dispatcher) — what `python -m qwen3_tts_cuda_graphs_b200.server --gpus N` runs behind
its HTTP endpoints.  All requests submitted at once; wall clock incl. prompt build, prefill, codec, D2H.
usage: python scripts/serving_box.py --gpus 4 --per-gpu 32 --concurrent 16"""
import argparse, json, os, sys, threading, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from qwen3_tts_cuda_graphs_b200 import server
from qwen3_tts_cuda_graphs_b200.serving import TTSRequest

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
ap.add_argument("--per-gpu", type=int, default=32, dest="per_gpu")
ap.add_argument("--concurrent", type=int, default=16)
ap.add_argument("--frames", type=int, default=250)
args = ap.parse_args()
backends = server.build_backends("synthetic://0.6B-Base", args.gpus, args.concurrent, 8, max_seq_len=1024)
disp = server.Dispatcher(backends)
ref_wav = bench.make_ref_wav()


def run(n):
    done = []
    def consume(h):
        a, sr = h.result()
        disp.release(h._backend_index)
        done.append(len(a) / sr)
    t0 = time.perf_counter()
    ths = []
    hs = []
    for i in range(n):
        h = disp.submit(TTSRequest(bench.TEXT + f" Request {i}.", ref_audio=ref_wav, ref_text=bench.REF_TEXT, language="English",
                                   max_new_tokens=args.frames, min_new_tokens=args.frames))
        hs.append(h)
        th = threading.Thread(target=consume, args=(h,), daemon=True)
        th.start(); ths.append(th)
    for th in ths:
        th.join(300)
    dt = time.perf_counter() - t0
    per = [0] * args.gpus
    for h in hs:
        per[h._backend_index] += 1
    return sum(done) / dt, dt, per, float(np.mean([h.ttfa_s for h in hs if h.ttfa_s is not None]) * 1000)


run(4 * args.gpus)
v, dt, per, ttfa = run(args.per_gpu * args.gpus)
print(json.dumps({"metric": "audio_seconds_per_second", "what": "serving.BatchScheduler per GPU behind server.Dispatcher, one process, all requests at once",
                  "n_gpus": args.gpus, "requests": args.per_gpu * args.gpus, "concurrent_per_gpu": args.concurrent, "frames": args.frames,
                  "value": round(v, 1), "seconds": round(dt, 3), "requests_per_gpu": per, "ttfa_ms_mean": round(ttfa, 1)}))
for b in backends:
    b.stop()
