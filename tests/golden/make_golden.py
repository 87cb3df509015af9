"""Generate the committed golden vectors from the REFERENCE's own code (run in the build container only).

* sampling_golden.pt — outputs of /root/reference/faster_qwen3_tts/sampling.py (loaded by file path, because
  importing the package pulls soundfile/qwen_tts which are absent) on seeded inputs: repetition penalty results,
  greedy ids, and the top-k / top-p candidate sets (found by probing the sampler with one-hot-ish draws).
* qwen3_stack_golden.pt — hidden states of transformers' Qwen3Model (the sibling of the un-vendored qwen_tts
  talker) on the `tiny` preset weights, fp32 and bf16, prefill + one cached decode step.

Nothing under tests/ reads /root/reference at run time; only this script does.
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_reference_sampling():
    spec = importlib.util.spec_from_file_location("ref_sampling", "/root/reference/faster_qwen3_tts/sampling.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sampling_vectors():
    ref = load_reference_sampling()
    g = torch.Generator().manual_seed(1234)
    cases = []
    for V, dtype in [(3072, torch.float32), (2048, torch.float32), (3072, torch.bfloat16), (257, torch.float32)]:
        for trial in range(3):
            logits = (3.0 * torch.randn(1, V, generator=g)).to(dtype)
            hist = torch.randint(0, V, (40,), generator=g)
            pen = ref.apply_repetition_penalty(logits.clone().unsqueeze(0), hist, 1.05)[0]
            smask = torch.zeros(V, dtype=torch.bool)
            smask[max(0, V - 1024):] = True
            eos = V - 900 if V > 1024 else 3
            smask[eos] = False
            greedy = ref.sample_logits(logits, temperature=0.9, top_k=50, top_p=1.0, do_sample=False,
                                       suppress_mask=smask, suppress_tokens=[eos])
            # candidate sets: a token survives iff the reference can still draw it -> draw many times
            sets = {}
            for (k, p) in [(50, 1.0), (5, 1.0), (0, 0.8), (20, 0.5), (50, 0.95)]:
                torch.manual_seed(trial)
                seen = torch.zeros(V, dtype=torch.bool)
                x = logits.float()
                for _ in range(400):
                    t = ref.sample_logits(x, temperature=0.9, top_k=k, top_p=p, do_sample=True, suppress_mask=smask)
                    seen[t] = True
                sets[(k, p)] = seen
            cases.append(dict(V=V, dtype=str(dtype), logits=logits, history=hist, penalised=pen, smask=smask, eos=eos,
                              greedy=greedy, drawn=sets))
    # the reference's own known-answer test (tests/test_sampling.py:10-21)
    logits = torch.zeros(1, 1, 10)
    logits[..., 7] = 1.0
    logits[..., 8] = -1.0
    others = [0, 1, 2, 3, 4, 5, 6, 8, 9]
    history = torch.tensor([7] + [others[i % len(others)] for i in range(1, 60)], dtype=torch.long)
    kat = ref.apply_repetition_penalty(logits.clone(), history, repetition_penalty=1.1)
    return dict(cases=cases, kat_logits=logits, kat_history=history, kat_out=kat)


def stack_vectors():
    from transformers import Qwen3Config, Qwen3Model

    from qwen3_tts_cuda_graphs_b200.config import preset
    from qwen3_tts_cuda_graphs_b200.weights import init_synthetic

    cfg = preset("tiny")
    t = cfg.talker
    w = init_synthetic(cfg, seed=3, norm_jitter=0.1, dtype=torch.float32, skip_text_embedding=True)
    hc = Qwen3Config(
        vocab_size=32, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
        num_hidden_layers=t.num_hidden_layers, num_attention_heads=t.num_attention_heads,
        num_key_value_heads=t.num_key_value_heads, head_dim=t.head_dim, rms_norm_eps=t.rms_norm_eps,
        rope_theta=t.rope_theta, attention_bias=False, max_position_embeddings=4096,
    )
    hc._attn_implementation = "eager"
    out = {}
    for dt in (torch.float32, torch.bfloat16):
        m = Qwen3Model(hc)
        inv = m.rotary_emb.inv_freq.clone()
        m = m.to(dt).eval()
        m.rotary_emb.inv_freq = inv  # from_pretrained keeps inv_freq in fp32; .to(bf16) would not
        sd = {k[len("talker.model."):]: v.to(dt) for k, v in w.items()
              if k.startswith("talker.model.layers") or k == "talker.model.norm.weight"}
        m.load_state_dict(sd, strict=False)
        x = torch.randn(1, 7, t.hidden_size, generator=torch.Generator().manual_seed(1)).to(dt)
        x2 = torch.randn(1, 1, t.hidden_size, generator=torch.Generator().manual_seed(2)).to(dt)
        with torch.no_grad():
            o = m(inputs_embeds=x, use_cache=True)
            o2 = m(inputs_embeds=x2, past_key_values=o.past_key_values, use_cache=True)
        out[str(dt)] = dict(x=x, x2=x2, prefill=o.last_hidden_state, step=o2.last_hidden_state)
    return out


if __name__ == "__main__":
    torch.save(sampling_vectors(), os.path.join(HERE, "sampling_golden.pt"))
    torch.save(stack_vectors(), os.path.join(HERE, "qwen3_stack_golden.pt"))
    print("wrote", os.listdir(HERE))
