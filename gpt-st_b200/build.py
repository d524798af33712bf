"""Build libgptst_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python gpt-st_b200/build.py [--force] [--tools]

No torch / pybind dependency: the library is a plain C-ABI shared object loaded with ctypes.
Objects are compiled in parallel and cached by source mtime under csrc/_build/.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libgptst_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def compile_one(src: str, force: bool) -> str:
    obj = os.path.join(OBJ, src[:-3] + ".o")
    sp = os.path.join(CSRC, src)
    if (not force and os.path.isfile(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(sp), headers_mtime())):
        return obj
    cmd = [NVCC, *FLAGS, "-c", sp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: compile_one(s, force), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, "-shared", "-cudart", "static", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[gptst_b200] built {LIB} from {len(srcs)} sources")
    return LIB


def build_tools(verbose: bool = True) -> None:
    """Stand-alone C++ checks under tools/ (kbench, htem_check): they dlopen the library, so they only need nvcc + libdl."""
    tools = os.path.join(os.path.dirname(HERE), "tools")
    for name in ("kbench", "htem_check", "cap_check"):
        src, exe = os.path.join(tools, name + ".cu"), os.path.join(tools, name)
        if os.path.isfile(exe) and os.path.getmtime(exe) >= os.path.getmtime(src):
            continue
        r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", exe, src, "-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for tools/{name}.cu:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[gptst_b200] built tools/{name}")


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    if "--tools" in sys.argv:
        build_tools()
