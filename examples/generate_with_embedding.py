"""Speech from a precomputed speaker embedding (x-vector mode) — the counterpart of the reference's
`examples/generate_with_embedding.py`, call for call: the embedding becomes a `voice_clone_prompt` dict, the prompt is built with
`_build_talker_inputs_local`, `fast_generate` runs the whole frame loop in one launch, `speech_tokenizer.decode` gives the waveform.
`tests/test_model_gpu.py::test_reference_example_flow_with_a_saved_speaker_embedding` holds this flow against
`generate_voice_clone(xvec_only=True)`.

    python examples/extract_speaker.py --ref_audio voice.wav --output speaker.pt
    python examples/generate_with_embedding.py --speaker speaker.pt --text "Hello world" --language English --output out.wav
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load_xvector_prompt(path: str, device: str = "cuda:0") -> dict:
    """A saved x-vector -> the voice_clone_prompt dict of the x-vector-only mode (no reference codes, no ICL)."""
    emb = torch.load(path, weights_only=True).to(device)
    return dict(ref_code=[None], ref_spk_embedding=[emb], x_vector_only_mode=[True], icl_mode=[False])


def main(argv=None):
    ap = argparse.ArgumentParser(description="TTS from a precomputed speaker embedding")
    ap.add_argument("--speaker", required=True, help=".pt file written by extract_speaker.py")
    ap.add_argument("--text", required=True)
    ap.add_argument("--language", default="Auto")
    ap.add_argument("--output", default="output.wav")
    ap.add_argument("--model_path", default="synthetic://0.6B-Base", help="checkpoint directory, cached hub id or synthetic://<preset>")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--max_new_tokens", type=int, default=2048)
    args = ap.parse_args(argv)

    from qwen3_tts_cuda_graphs_b200 import FasterQwen3TTS
    from qwen3_tts_cuda_graphs_b200.cli import write_audio
    from qwen3_tts_cuda_graphs_b200.generate import fast_generate

    print(f"Loading model from {args.model_path}...")
    model = FasterQwen3TTS.from_pretrained(args.model_path, device="cuda", dtype=torch.bfloat16)
    vcp = load_xvector_prompt(args.speaker, device=args.device)
    input_ids = model.model._tokenize_texts([model.model._build_assistant_text(args.text)])
    tie, tam, tth, tpe = model._build_talker_inputs_local(
        m=model.model.model, input_ids=input_ids, ref_ids=[None], voice_clone_prompt=vcp, languages=[args.language], speakers=None,
        non_streaming_mode=False)
    print(f"Prefill length: {tie.shape[1]} tokens")
    model._warmup(tie.shape[1])  # nothing is captured lazily here; kept because the reference's callers do it

    talker = model.model.model.talker
    config = model.model.model.config.talker_config
    talker.rope_deltas = None
    codec_ids, timing = fast_generate(talker, tie, tam, tth, tpe, config, model.predictor_graph, model.talker_graph,
                                      temperature=0.9, top_k=50, do_sample=True, max_new_tokens=args.max_new_tokens)
    if codec_ids is None or codec_ids.numel() == 0:
        print("ERROR: generation returned no tokens")
        sys.exit(1)
    wavs, sr = model.model.speech_tokenizer.decode([{"audio_codes": codec_ids.to(model.device)}])
    audio = wavs[0].flatten().float().cpu().numpy()
    write_audio(args.output, audio, sr)
    n = timing["steps"]
    gen_s = timing["prefill_ms"] / 1000 + timing["decode_s"]
    print(f"Saved {args.output} ({n * 0.08:.1f}s audio, {gen_s:.2f}s gen, RTF {n * 0.08 / gen_s:.2f})")
    print(f"  Prefill: {timing['prefill_ms']:.0f}ms | Decode: {n} steps @ {timing['ms_per_step']:.2f}ms/step")


if __name__ == "__main__":
    main()
