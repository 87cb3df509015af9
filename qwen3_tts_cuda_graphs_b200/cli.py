"""Command line (reference: `faster_qwen3_tts/cli.py` — same sub-commands and flags):

    clone   --model M --text T --language L --output out.wav --ref-audio a.wav --ref-text "..." [--xvec-only] [--streaming]
    custom  --model M --text T --output out.wav --speaker aiden [--instruct "..."] [--list-speakers]
    design  --model M --text T --output out.wav --instruct "..."
    serve   --mode clone|custom|design --model M ...      one utterance per stdin line, model kept hot (cli.py:186-305)
    http    ...                                            the OpenAI-compatible server (server.py)

Beyond the reference: `serve --concurrency N` hands the stdin lines to the continuous-batching scheduler, so N utterances
decode in lock-step instead of one after the other.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import wave

import numpy as np


def _load_model(model_id: str, device: str, dtype: str, max_streams: int = 1):
    import torch

    from .model import FasterQwen3TTS

    torch_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    return FasterQwen3TTS.from_pretrained(model_id, device=device, dtype=torch_dtype, attn_implementation="sdpa", max_seq_len=2048,
                                          max_streams=max_streams)


def write_audio(out_path: str, audio: np.ndarray, sr: int):
    """16-bit PCM WAV (cli.py:31-33 writes through soundfile, which is not in this image)."""
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    pcm = np.clip(np.asarray(audio, dtype=np.float32) * 32768.0, -32768, 32767).astype("<i2")
    with wave.open(out_path, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(int(sr))
        wf.writeframes(pcm.tobytes())


def _stream_to_audio(gen):
    chunks, sr = [], None
    for audio_chunk, sr, _ in gen:
        chunks.append(audio_chunk)
    if not chunks:
        return np.zeros(1, dtype=np.float32), 24000
    return np.concatenate(chunks), sr


def _sampling(args) -> dict:
    return dict(max_new_tokens=args.max_new_tokens, temperature=args.temperature, top_k=args.top_k, do_sample=not args.greedy,
                repetition_penalty=args.repetition_penalty)


def _call(model, mode: str, text: str, args):
    """(method name stem, keyword arguments) of the reference API for one utterance in `mode`."""
    kw = dict(text=text, language=args.language, **_sampling(args))
    if mode == "clone":
        kw.update(ref_audio=args.ref_audio, ref_text=args.ref_text, xvec_only=bool(getattr(args, "xvec_only", False)),
                  non_streaming_mode=args.non_streaming_mode)
        stem = "generate_voice_clone"
    elif mode == "custom":
        kw.update(speaker=args.speaker, instruct=args.instruct)
        stem = "generate_custom_voice"
    else:
        kw.update(instruct=args.instruct)
        stem = "generate_voice_design"
    return stem, kw


def _generate(model, mode: str, text: str, args):
    stem, kw = _call(model, mode, text, args)
    if args.streaming:
        return _stream_to_audio(getattr(model, stem + "_streaming")(chunk_size=args.chunk_size, **kw))
    audio_list, sr = getattr(model, stem)(**kw)
    return audio_list[0], sr


def _report(out_path, audio, sr, start):
    total = time.perf_counter() - start
    dur = len(audio) / sr if sr else 0.0
    print(f"Wrote {out_path} (dur {dur:.2f}s, RTF {dur / total if total > 0 else 0.0:.2f})")


def _one_shot(mode):
    def run(args):
        model = _load_model(args.model, args.device, args.dtype)
        if mode == "custom" and args.list_speakers:
            for s in model.model.get_supported_speakers():
                print(s)
            return
        if mode == "custom" and not args.speaker:
            print("ERROR: --speaker is required (or --list-speakers)")
            sys.exit(2)
        start = time.perf_counter()
        audio, sr = _generate(model, mode, args.text, args)
        write_audio(args.output, audio, sr)
        _report(args.output, audio, sr, start)
    return run


def cmd_serve(args):
    if args.mode == "clone" and not (args.ref_audio and args.ref_text is not None):
        print("ERROR: --ref-audio and --ref-text are required for clone mode")
        sys.exit(2)
    if args.mode == "custom" and not args.speaker:
        print("ERROR: --speaker is required for custom mode")
        sys.exit(2)
    if args.mode == "design" and not args.instruct:
        print("ERROR: --instruct is required for design mode")
        sys.exit(2)
    model = _load_model(args.model, args.device, args.dtype, max_streams=max(1, args.concurrency))
    print("Server started. Enter text per line. Type 'exit' or 'quit' to stop.")
    sched = None
    if args.concurrency > 1:
        from .serving import BatchScheduler

        sched = BatchScheduler(model, chunk_frames=args.chunk_size, max_concurrent=args.concurrency).start()
    waiting = []
    idx = 1

    def drain(block_all: bool):
        while waiting and (block_all or len(waiting) >= 4 * args.concurrency):
            out_path, h, start = waiting.pop(0)
            try:
                audio, sr = h.result()
            except Exception as e:  # one bad line must not lose the utterances decoding beside it
                print(f"ERROR: {out_path}: {e}", file=sys.stderr)
                continue
            write_audio(out_path, audio, sr)
            _report(out_path, audio, sr, start)

    try:
        for line in sys.stdin:
            text = line.strip()
            if not text:
                continue
            if text.lower() in ("exit", "quit", "stop"):
                break
            out_path = os.path.join(args.output_dir, f"out_{idx:04d}.wav")
            idx += 1
            start = time.perf_counter()
            if sched is None:
                audio, sr = _generate(model, args.mode, text, args)
                write_audio(out_path, audio, sr)
                _report(out_path, audio, sr, start)
                continue
            from .serving import TTSRequest

            _, kw = _call(model, args.mode, text, args)
            kw["kind"] = {"clone": "voice_clone", "custom": "custom_voice", "design": "voice_design"}[args.mode]
            waiting.append((out_path, sched.submit(TTSRequest(**kw)), start))
            drain(False)
        drain(True)
    finally:
        if sched is not None:
            sched.stop()


def cmd_http(args):
    from . import server

    server.main(args.rest)


def build_parser():
    p = argparse.ArgumentParser(prog="faster-qwen3-tts", description="Qwen3-TTS on the fq3 B200 engine")
    p.add_argument("--device", default="cuda", help="Device (cuda)")
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"], help="Model dtype (the engine computes in bf16)")
    sub = p.add_subparsers(dest="cmd", required=True)

    def add_sampling(sp):
        sp.add_argument("--max-new-tokens", type=int, default=2048)
        sp.add_argument("--temperature", type=float, default=0.9)
        sp.add_argument("--top-k", type=int, default=50)
        sp.add_argument("--repetition-penalty", type=float, default=1.05)
        sp.add_argument("--greedy", action="store_true", help="Disable sampling")
        sp.add_argument("--streaming", action="store_true", help="Use streaming generation")
        g = sp.add_mutually_exclusive_group()
        g.add_argument("--non-streaming-mode", dest="non_streaming_mode", action="store_true", help="Full text in the prefill (default)")
        g.add_argument("--no-non-streaming-mode", dest="non_streaming_mode", action="store_false", help="Feed text step by step")
        sp.set_defaults(non_streaming_mode=True)
        sp.add_argument("--chunk-size", type=int, default=8, help="Streaming chunk size")

    def add_common(sp):
        sp.add_argument("--text", required=True, help="Text to synthesize")
        sp.add_argument("--language", default="Auto", help="Language (Auto, English, French, ...)")
        sp.add_argument("--output", required=True, help="Output wav path")
        sp.add_argument("--model", required=True, help="Checkpoint directory, cached hub id or synthetic://<preset>")
        add_sampling(sp)

    sp = sub.add_parser("clone", help="Voice cloning (reference audio)")
    add_common(sp)
    sp.add_argument("--ref-audio", required=True, help="Reference audio path")
    sp.add_argument("--ref-text", required=True, help="Reference transcript")
    sp.add_argument("--xvec-only", action="store_true", help="Use speaker embedding only")
    sp.set_defaults(fn=_one_shot("clone"))

    sp = sub.add_parser("custom", help="CustomVoice model (speaker IDs)")
    add_common(sp)
    sp.add_argument("--speaker", help="Speaker ID")
    sp.add_argument("--instruct", default="", help="Optional instruction")
    sp.add_argument("--list-speakers", action="store_true", help="List available speaker IDs")
    sp.set_defaults(fn=_one_shot("custom"))

    sp = sub.add_parser("design", help="VoiceDesign model (instruction-based)")
    add_common(sp)
    sp.add_argument("--instruct", required=True, help="Voice/style instruction")
    sp.set_defaults(fn=_one_shot("design"))

    sp = sub.add_parser("serve", help="Keep model hot and generate multiple requests from stdin")
    sp.add_argument("--mode", required=True, choices=["clone", "custom", "design"])
    sp.add_argument("--model", required=True, help="Checkpoint directory, cached hub id or synthetic://<preset>")
    sp.add_argument("--language", default="Auto", help="Language (Auto, English, French, ...)")
    sp.add_argument("--ref-audio", help="Reference audio path (clone)")
    sp.add_argument("--ref-text", help="Reference transcript (clone)")
    sp.add_argument("--speaker", help="Speaker ID (custom)")
    sp.add_argument("--instruct", default="", help="Instruction (custom/design)")
    add_sampling(sp)
    sp.add_argument("--output-dir", default="outputs", help="Directory for output wavs")
    sp.add_argument("--concurrency", type=int, default=1, help="utterances decoded in lock-step (continuous batching)")
    sp.set_defaults(fn=cmd_serve)

    sp = sub.add_parser("http", help="OpenAI-compatible HTTP server (see server.py --help)")
    sp.add_argument("rest", nargs=argparse.REMAINDER)
    sp.set_defaults(fn=cmd_http)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    args.fn(args)


if __name__ == "__main__":
    main()
