"""Host build of a CUDA-core kernel source under the SIMT emulator (``cuda_emu.h``).

``build("layer_ops.cu")`` rewrites every ``kernel<<<grid, block, smem, stream>>>(args);`` into
``flpr_emu::launch(grid, block, stream, [=]() { kernel(args); });``, turns ``__shared__`` declarations into storage owned
by the running block, removes the functions whose body is inline PTX (the emulator header supplies host versions),
swaps ``#include "ptx.cuh"`` for the emulator header, compiles the result with ``g++`` (no nvcc, no CUDA runtime call)
and returns the path of a shared library that exports the same ``extern "C"`` entry points as the real one - so the
ctypes wrappers of ``flpr_b200.ops`` (``tests/test_cpu_emulated_kernels.py``) or a multi-rank harness
(``comm_harness.py``, ``tests/test_cpu_emulated_collectives.py``) can drive it with CPU tensors."""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import tempfile
from typing import List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "federated-lifelong-person-reid_b200", "csrc")
CUDA_INCLUDE = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


def _match_forward(s: str, i: int, open_ch: str, close_ch: str) -> int:
    """``s[i] == open_ch``: index just past the matching ``close_ch``."""
    assert s[i] == open_ch, (s[i - 20:i + 20], open_ch)
    depth = 0
    while i < len(s):
        if s[i] == open_ch:
            depth += 1
        elif s[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced " + open_ch)


def _split_top_level(s: str) -> List[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _kernel_name_start(s: str, end: int) -> int:
    """``s[:end]`` ends with a kernel name, optionally with template arguments (``ln_rows_kernel<4>``): its start."""
    i = end
    if s[i - 1] == ">":                                   # template arguments
        depth = 0
        while i > 0:
            i -= 1
            if s[i] == ">":
                depth += 1
            elif s[i] == "<":
                depth -= 1
                if depth == 0:
                    break
    while i > 0 and (s[i - 1].isalnum() or s[i - 1] in "_:"):
        i -= 1
    return i


def rewrite_launches(src: str) -> Tuple[str, int]:
    out, pos, n = [], 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        name_start = _kernel_name_start(src, k)
        cfg_end = src.index(">>>", k)
        cfg = _split_top_level(src[k + 3:cfg_end])
        assert 2 <= len(cfg) <= 4, cfg
        args_open = cfg_end + 3
        while src[args_open].isspace():
            args_open += 1
        args_end = _match_forward(src, args_open, "(", ")")
        name = src[name_start:k].strip()
        args = src[args_open + 1:args_end - 1]
        out.append(src[pos:name_start])
        smem = f"(size_t)({cfg[2]})" if len(cfg) >= 3 else "0"
        key = f"(const void*)({cfg[3]})" if len(cfg) >= 4 else "nullptr"
        out.append(f"flpr_emu::launch(flpr_emu::dims({cfg[0]}), flpr_emu::dims({cfg[1]}), {smem}, {key}, "
                   f"[=]() {{ {name}({args}); }})")
        pos = args_end
        n += 1
    return "".join(out), n


_DYN_SHARED_RE = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w:]*)\s+(\w+)\s*\[\s*\]\s*;")
_SHARED_LIST_RE = re.compile(r"__shared__\s+([A-Za-z_][\w:]*)\s+([^;()=]*,[^;()=]*);")
_SHARED_RE = re.compile(r"__shared__\s+([A-Za-z_][\w:]*(?:\s*<[^;<>]*>)?)\s+(\w+)\s*((?:\[[^\]]*\]\s*)*);")


def rewrite_shared(src: str) -> Tuple[str, int]:
    """``__shared__ T name[a][b];`` -> a reference to storage owned by the running block (``flpr_emu::shared``)."""
    n = 0
    src = _DYN_SHARED_RE.sub(lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>("
                                       f"flpr_emu::dyn_shared());", src)

    def one(ty, name, dims):
        nonlocal n
        n += 1
        return (f"using flpr_sh_t{n} = {ty}{dims}; "
                f"flpr_sh_t{n}& {name} = *flpr_emu::shared<flpr_sh_t{n}>({n});")

    def repl(m):
        return one(m.group(1), m.group(2), m.group(3).replace(" ", ""))

    def repl_list(m):                                    # ``__shared__ int a, b[4], c;``
        out = []
        for decl in m.group(2).split(","):
            d = re.match(r"\s*(\w+)\s*((?:\[[^\]]*\]\s*)*)$", decl)
            assert d, decl
            out.append(one(m.group(1), d.group(1), d.group(2).replace(" ", "")))
        return " ".join(out)

    src = _SHARED_LIST_RE.sub(repl_list, src)
    src = _SHARED_RE.sub(repl, src)
    code = re.sub(r"//[^\n]*", "", src)
    assert "__shared__" not in code, "an unhandled __shared__ declaration would silently become a per-thread local"
    return src, n


def strip_inline_ptx_functions(src: str) -> Tuple[str, List[str]]:
    """Remove every function whose body is inline PTX (``asm volatile``): the emulator header supplies host versions of
    the same names (``gtimer``, ``multimem_*``). Returns the new source and the names removed."""
    removed = []
    while True:
        k = src.find("asm volatile")
        if k < 0:
            return src, removed
        start = src.rfind("\n__device__", 0, k)
        assert start >= 0, "inline PTX outside a __device__ function"
        start += 1
        brace = src.index("{", start)
        assert brace < k
        end = _match_forward(src, brace, "{", "}")
        header = src[start:brace]
        removed.append(re.search(r"(\w+)\s*\([^()]*\)\s*$", header.strip()).group(1))
        src = src[:start] + src[end:]


def emulated_source(cu_name: str, mutate=None) -> str:
    with open(os.path.join(CSRC, cu_name)) as f:
        src = f.read()
    if mutate is not None:                       # mutation checks: a deliberately broken kernel the tests must notice
        src = mutate(src)
    assert '#include "ptx.cuh"' in src
    src = src.replace('#include "ptx.cuh"', '#include "cuda_emu.h"')
    src, removed = strip_inline_ptx_functions(src)
    assert set(removed) <= {"gtimer", "multimem_ld_reduce_add_f4", "multimem_st_f4"}, removed
    if mutate is None:
        assert '#include "cuda_emu.h"' in src
    src, _ = rewrite_shared(src)
    src, n = rewrite_launches(src)
    assert n > 0 and "<<<" not in src
    return src


def build(cu_name: str = "layer_ops.cu", mutate=None) -> str:
    """Returns the path of the emulated shared library (cached per source hash in the temp dir)."""
    src = emulated_source(cu_name, mutate)
    with open(os.path.join(HERE, "cuda_emu.h")) as f:
        hdr = f.read()
    tag = hashlib.sha256((src + hdr).encode()).hexdigest()[:16]
    out_dir = os.path.join(tempfile.gettempdir(), f"flpr_emu_{tag}")
    lib = os.path.join(out_dir, "lib" + os.path.splitext(cu_name)[0] + "_emu.so")
    if os.path.exists(lib):
        return lib
    gxx = shutil.which("g++")
    if gxx is None or not os.path.isdir(CUDA_INCLUDE):
        raise RuntimeError("g++ / CUDA headers not available")
    os.makedirs(out_dir, exist_ok=True)
    cpp = os.path.join(out_dir, os.path.splitext(cu_name)[0] + "_emu.cpp")
    with open(cpp, "w") as f:
        f.write(src)
    tmp = lib + f".{os.getpid()}.tmp"
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wno-attributes", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-I", HERE, "-I", CSRC, "-I", CUDA_INCLUDE, cpp, "-o", tmp]
    cudart = os.path.join(os.path.dirname(CUDA_INCLUDE), "lib64")
    if os.path.exists(os.path.join(cudart, "libcudart.so")):      # host-side helpers of fedcomm.cu reference the runtime
        cmd += ["-L", cudart, "-lcudart", f"-Wl,-rpath,{cudart}"]   # (never called under emulation)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + r.stderr[-4000:])
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    print(build())
