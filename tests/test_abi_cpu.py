"""The C-ABI libraries load and export every symbol include/*.h declares (no compute without a GPU)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = {"fq3.h": "libfq3.so", "fq3_codec.h": "libfq3codec.so"}


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fq3[a-z0-9_]*)\s*\(", src)))


@pytest.mark.parametrize("header", [h for h in HEADERS if os.path.exists(os.path.join(ROOT, "include", h))])
def test_library_exports_every_declared_symbol(header):
    from qwen3_tts_cuda_graphs_b200 import build
    path = build.build_lib(HEADERS[header])
    assert os.path.exists(path), f"{path} not built"
    lib = ctypes.CDLL(path)
    names = declared(header)
    assert len(names) >= 4
    for n in names:
        assert hasattr(lib, n), f"{HEADERS[header]} does not export {n}"


def test_ctypes_table_covers_header():
    from qwen3_tts_cuda_graphs_b200 import _lib
    assert set(declared("fq3.h")) == set(_lib.SYMBOLS)


def test_abi_version_without_gpu():
    from qwen3_tts_cuda_graphs_b200 import _lib
    lib = _lib.load()
    assert lib.fq3_abi_version() == 2


def test_engine_refuses_cpu_arena():
    import torch
    from qwen3_tts_cuda_graphs_b200.config import preset
    from qwen3_tts_cuda_graphs_b200.engine import Engine
    from qwen3_tts_cuda_graphs_b200.weights import init_synthetic, pack_arena
    cfg = preset("tiny")
    w = init_synthetic(cfg, skip_text_embedding=True)
    arena = pack_arena(cfg, w, 64, torch.device("cpu"))
    with pytest.raises(ValueError):
        Engine(cfg, arena, max_seq_len=64)


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "qwen3_tts_cuda_graphs_b200")
    for f in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True):
        src = open(f).read()
        assert "import oracle" not in src and "from oracle" not in src, f
