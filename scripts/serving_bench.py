"""Serving-layer throughput and latency on one GPU (BASELINE configs[4] through serving.BatchScheduler, not through a prepared
batch): `--requests` voice-clone requests of `--frames` frames arrive open-loop at `--rate` requests/s (0 = all at once) and
are decoded by the continuous-batching scheduler with up to `--concurrent` lock-step streams.  Reports aggregate
audio-seconds/second (wall clock, prompt build + prefill + decode + codec + D2H included), time to first audio per request
(mean / p50 / p95, queueing included) and how full the launches were.
usage: python scripts/serving_bench.py --concurrent 16 --requests 64 --frames 250 [--rate 20]"""
import argparse, json, os, sys, threading, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
from qwen3_tts_cuda_graphs_b200.serving import BatchScheduler, TTSRequest

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="synthetic://0.6B-Base")
ap.add_argument("--concurrent", type=int, default=16)
ap.add_argument("--requests", type=int, default=64)
ap.add_argument("--frames", type=int, default=250)
ap.add_argument("--rate", type=float, default=0.0)
ap.add_argument("--chunk", type=int, default=8)
ap.add_argument("--lanes", type=int, default=4)
ap.add_argument("--codec", default="auto")
ap.add_argument("--overlap", default="auto", help="auto | host | gpu | none")
ap.add_argument("--emit-every", type=int, default=1, dest="emit_every")
args = ap.parse_args()

model = FasterQwen3TTS.from_pretrained(args.model, device="cuda", dtype=torch.bfloat16, max_seq_len=1024, seed=0, max_streams=args.concurrent)
ref_wav = bench.make_ref_wav()
texts = [bench.TEXT + f" Request number {i}." for i in range(max(args.requests, 4))]


def run(n_req, rate):
    handles, done = [], []
    with BatchScheduler(model, chunk_frames=args.chunk, max_concurrent=args.concurrent, codec_lanes=args.lanes, codec_mode=args.codec, overlap_codec=args.overlap, emit_every=args.emit_every) as sched:
        def consume(h):
            a, sr = h.result()
            done.append((h, len(a) / sr))
        t0 = time.perf_counter()
        threads = []
        for i in range(n_req):
            if rate > 0:
                time.sleep(max(0.0, t0 + i / rate - time.perf_counter()))
            h = sched.submit(TTSRequest(texts[i], ref_audio=ref_wav, ref_text=bench.REF_TEXT, language="English", max_new_tokens=args.frames,
                                        min_new_tokens=args.frames))
            th = threading.Thread(target=consume, args=(h,))
            th.start()
            threads.append(th)
            handles.append(h)
        for th in threads:
            th.join()
        dt = time.perf_counter() - t0
        stats = dict(sched.stats)
    ttfa = np.array([h.ttfa_s for h, _ in done]) * 1000
    lat = np.array([h.t_done - h.t_submit for h, _ in done]) * 1000
    audio = sum(s for _, s in done)
    return {"requests": n_req, "rate_per_s": rate, "concurrent": args.concurrent, "frames": args.frames, "chunk_frames": args.chunk, "codec_lanes": args.lanes, "codec_mode": sched.codec_mode, "overlap_codec": args.overlap, "emit_every": args.emit_every,
            "audio_s_per_s": round(audio / dt, 1), "seconds": round(dt, 3),
            "ttfa_ms": {"mean": round(float(ttfa.mean()), 1), "p50": round(float(np.percentile(ttfa, 50)), 1), "p95": round(float(np.percentile(ttfa, 95)), 1)},
            "latency_ms": {"mean": round(float(lat.mean()), 1), "p95": round(float(np.percentile(lat, 95)), 1)},
            "loop_seconds": {k[2:]: round(v, 3) for k, v in stats.items() if k.startswith("t_")}, "launches": stats["launches"], "mean_streams_per_launch": round(stats["frames"] / args.chunk / max(stats["launches"], 1), 2)}


run(min(args.concurrent, 4), 0.0)  # warm-up: plans, graphs, voice prompt cache
print(json.dumps(run(args.requests, args.rate)))
