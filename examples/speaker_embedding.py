"""A voice as a 4 KB file: x-vector mode end to end on this package.

Two steps, one script (the reference ships them as `examples/extract_speaker.py` and `examples/generate_with_embedding.py`; the
calls below are the ones those two make, against this package's wrapper instead of `qwen_tts`):

    python examples/speaker_embedding.py extract --ref-audio voice.wav --out speaker.pt
    python examples/speaker_embedding.py speak   --speaker speaker.pt --text "Hello world" --language English --out hello.wav

`extract` encodes the clip once (`create_voice_clone_prompt(..., x_vector_only_mode=True)`) and saves the embedding.
`speak` turns the saved embedding into a `voice_clone_prompt` dict, builds the prompt rows with `_build_talker_inputs_local`
(10 rows + text instead of the ~80+ of an ICL prompt: the shortest prefill, no accent carried over from the clip), runs the whole
frame loop in one `fast_generate` call and decodes with `speech_tokenizer.decode`.
`tests/test_model_gpu.py::test_reference_example_flow_with_a_saved_speaker_embedding` holds exactly this flow against
`generate_voice_clone(xvec_only=True)` on the GPU.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DEFAULT_MODEL = "synthetic://0.6B-Base"  # or a checkpoint directory / a cached hub id


def extract(args) -> None:
    from qwen3_tts_cuda_graphs_b200.base_model import Qwen3TTSBaseModel

    wrapper = Qwen3TTSBaseModel.from_pretrained(args.model, device_map=args.device, torch_dtype=torch.bfloat16)
    (item,) = wrapper.create_voice_clone_prompt(ref_audio=args.ref_audio, ref_text="", x_vector_only_mode=True)
    vec = item.ref_spk_embedding.detach().cpu()
    torch.save(vec, args.out)
    print(f"{args.out}: speaker embedding {tuple(vec.shape)} {vec.dtype}, {vec.numel() * vec.element_size()} bytes")


def prompt_from_file(path: str, device: str) -> dict:
    """The voice_clone_prompt of the x-vector-only mode: one embedding, no reference codes, no in-context part."""
    vec = torch.load(path, weights_only=True).to(device)
    return {"ref_spk_embedding": [vec], "ref_code": [None], "x_vector_only_mode": [True], "icl_mode": [False]}


def speak(args) -> None:
    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
    from qwen3_tts_cuda_graphs_b200.cli import write_audio
    from qwen3_tts_cuda_graphs_b200.generate import fast_generate

    tts = FasterQwen3TTS.from_pretrained(args.model, device="cuda", dtype=torch.bfloat16)
    wrapper, inner = tts.model, tts.model.model
    ids = wrapper._tokenize_texts([wrapper._build_assistant_text(args.text)])
    rows, mask, trailing, pad_row = tts._build_talker_inputs_local(
        m=inner, input_ids=ids, ref_ids=[None], voice_clone_prompt=prompt_from_file(args.speaker, args.device),
        languages=[args.language], speakers=None, non_streaming_mode=False)
    tts._warmup(rows.shape[1])  # a no-op beyond bookkeeping here (nothing is captured lazily); the reference's callers do it
    inner.talker.rope_deltas = None
    codes, timing = fast_generate(inner.talker, rows, mask, trailing, pad_row, inner.config.talker_config, tts.predictor_graph,
                                  tts.talker_graph, max_new_tokens=args.max_new_tokens, do_sample=not args.greedy, temperature=0.9, top_k=50)
    if codes is None:
        sys.exit("generation returned no frames")
    (wav,), rate = wrapper.speech_tokenizer.decode([{"audio_codes": codes.to(tts.device)}])
    write_audio(args.out, wav.flatten().float().cpu().numpy(), rate)
    seconds, spent = timing["steps"] * 0.08, timing["prefill_ms"] / 1000 + timing["decode_s"]
    print(f"{args.out}: {seconds:.1f} s of audio in {spent:.2f} s (RTF {seconds / spent:.1f}); prompt {rows.shape[1]} rows, "
          f"prefill {timing['prefill_ms']:.1f} ms, {timing['steps']} frames at {timing['ms_per_step']:.2f} ms")


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--model", default=DEFAULT_MODEL)
    ap.add_argument("--device", default="cuda:0")
    sub = ap.add_subparsers(dest="step", required=True)
    ex = sub.add_parser("extract", help="reference clip -> speaker.pt")
    ex.add_argument("--ref-audio", required=True)
    ex.add_argument("--out", required=True)
    ex.set_defaults(run=extract)
    sp = sub.add_parser("speak", help="speaker.pt + text -> wav")
    sp.add_argument("--speaker", required=True)
    sp.add_argument("--text", required=True)
    sp.add_argument("--language", default="Auto")
    sp.add_argument("--out", default="output.wav")
    sp.add_argument("--max-new-tokens", type=int, default=2048)
    sp.add_argument("--greedy", action="store_true")
    sp.set_defaults(run=speak)
    args = ap.parse_args(argv)
    args.run(args)


if __name__ == "__main__":
    main()
