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


def test_ctypes_struct_layouts_match_the_headers(tmp_path):
    """The ctypes mirrors (codec.Op, _lib.Policy / SubPolicy / Status / StackDesc / ModelDesc) have the size and the field offsets
    gcc gives the C structs of include/*.h — a silent layout drift would corrupt every call."""
    import shutil
    import subprocess

    from qwen3_tts_cuda_graphs_b200 import _lib
    from qwen3_tts_cuda_graphs_b200 import codec

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = [("fq3c_op", codec.Op, "fq3_codec.h"), ("fq3_policy", _lib.Policy, "fq3.h"), ("fq3_subpolicy", _lib.SubPolicy, "fq3.h"),
             ("fq3_status", _lib.Status, "fq3.h"), ("fq3_stack_desc", _lib.StackDesc, "fq3.h"), ("fq3_model_desc", _lib.ModelDesc, "fq3.h")]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fq3.h"', '#include "fq3_codec.h"', "int main(void) {"]
    for cname, _, hdr in pairs:
        src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", hdr)).read(), flags=re.S)
        body = re.search(r"typedef struct " + cname + r"\s*\{(.*?)\}\s*" + cname + r"\s*;", src, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                name = re.sub(r"\[.*?\]", "", part.strip().split()[-1].lstrip("*"))
                fields.append(name)
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for line in out.splitlines():
        cname, f, v = line.split()
        got.setdefault(cname, {})[f] = int(v)
    for cname, ctype, _ in pairs:
        assert ctypes.sizeof(ctype) == got[cname].pop("size"), cname
        mirror = {n: getattr(ctype, n).offset for n, *_ in ctype._fields_}
        assert list(mirror.values()) == list(got[cname].values()), (cname, mirror, got[cname])
