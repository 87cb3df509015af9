"""Request-parallel throughput of the whole box (BASELINE configs[4]): every rank (one per GPU, torchrun) decodes `--streams`
utterances in lock-step groups on its own replica; aggregate audio-seconds per second = sum over ranks / max time over ranks.
No collective on the data path (replicas only); the barrier and the max-reduction are for timing.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/batched_scale.py --streams 16"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
from qwen3_tts_cuda_graphs_b200.generate import fast_generate_batch

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=16)
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = f"cuda:{local}"
model = FasterQwen3TTS.from_pretrained("synthetic://0.6B-Base", device=dev, dtype=torch.bfloat16, attn_implementation="eager", max_seq_len=1024,
                                       seed=0, max_streams=args.streams)
ref_wav = bench.make_ref_wav()
m, _, _, tie, tam, tth, tpe, _ = model._prepare_generation(bench.TEXT, ref_wav, bench.REF_TEXT, language="English", non_streaming_mode=True)
reqs = [(tie, tam, tth, tpe)] * args.streams

def step():
    codes, _ = fast_generate_batch(model.talker_graph, model.predictor_graph, reqs, max_new_tokens=args.frames, min_new_tokens=args.frames)
    n = 0
    for c in codes:
        a, sr = model._decode_full(m, c)
        n += len(a[0])
    return n / sr

step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
audio = sum(step() for _ in range(args.steps))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt, audio], dtype=torch.float64, device=dev)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    dt, audio = float(tmax[0]), float(tsum[1])
if rank == 0:
    print(json.dumps({"metric": "audio_seconds_per_second", "what": f"request-parallel, {args.streams} streams per GPU in lock-step groups of up to "
                      f"{model.model.engine.lockstep_group}, {args.frames} frames per utterance, prefill and codec included, replicas only",
                      "n_gpus": world, "streams_per_gpu": args.streams, "value": audio / dt, "seconds": dt, "steps": args.steps}))
if world > 1:
    dist.destroy_process_group()
