"""Save a voice's speaker embedding (x-vector) to a .pt file — the counterpart of the reference's `examples/extract_speaker.py`,
on this package's model wrapper (no `qwen_tts` import): load once, encode the clip with `create_voice_clone_prompt(x_vector_only_mode=True)`,
`torch.save` the vector.  `examples/generate_with_embedding.py` consumes the file.

    python examples/extract_speaker.py --ref_audio voice.wav --output speaker.pt [--model_path <dir | synthetic://0.6B-Base>]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser(description="Extract a speaker embedding from reference audio")
    ap.add_argument("--ref_audio", required=True, help="reference clip (wav)")
    ap.add_argument("--output", required=True, help="where the embedding goes (.pt)")
    ap.add_argument("--model_path", default="synthetic://0.6B-Base", help="checkpoint directory, cached hub id or synthetic://<preset>")
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)

    from qwen3_tts_cuda_graphs_b200.base_model import Qwen3TTSBaseModel

    print(f"Loading model from {args.model_path}...")
    model = Qwen3TTSBaseModel.from_pretrained(args.model_path, device_map=args.device, torch_dtype=torch.bfloat16)
    print(f"Extracting speaker embedding from {args.ref_audio}...")
    items = model.create_voice_clone_prompt(ref_audio=args.ref_audio, ref_text="", x_vector_only_mode=True)
    emb = items[0].ref_spk_embedding.cpu()
    torch.save(emb, args.output)
    print(f"Saved speaker embedding to {args.output}")
    print(f"  Shape: {tuple(emb.shape)}, dtype: {emb.dtype}, size: {emb.nelement() * emb.element_size()} bytes")


if __name__ == "__main__":
    main()
