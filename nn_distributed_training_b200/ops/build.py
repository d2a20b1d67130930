"""In-tree build of the sm_100a extension (``ops/_C*.so``).

Kernels are compiled directly with nvcc for ``-gencode arch=compute_100a,code=sm_100a
-lineinfo`` (torch's arch list is bypassed: tcgen05/TMA need the ``a`` target); only the
binding translation units see the PyTorch headers, so kernel edits rebuild in seconds.
Objects are cached under ``ops/build/`` keyed by source+flags hash.  The .so stays in the
tree (git-ignored) so it travels with a gpurun snapshot.

    python -m nn_distributed_training_b200.ops.build [--force] [-v]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
import sysconfig
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
              "--use_fast_math"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wno-deprecated-declarations"]


def _cuda_home() -> str:
    for cand in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if cand and os.path.exists(os.path.join(cand, "bin", "nvcc")):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    return cu, cpp


def _hash(path: str, flags: List[str]) -> str:
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    for f in sorted(os.listdir(CSRC)):  # headers affect every TU; cheap to hash all
        if f.endswith((".h", ".cuh")) or os.path.join(CSRC, f) == path:
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


def _run(cmd: List[str], verbose: bool):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def so_path() -> str:
    return os.path.join(HERE, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False, verbose: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce

    cuda = _cuda_home()
    nvcc = os.path.join(cuda, "bin", "nvcc")
    os.makedirs(BUILD, exist_ok=True)
    cu, cpp = _sources()
    inc = [f"-I{CSRC}", f"-I{os.path.join(cuda, 'include')}"]
    torch_inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    abi = [f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_EXTENSION_NAME=_C",
           "-DTORCH_API_INCLUDE_EXTENSION_H"]
    try:
        import pybind11
        pyb = [f"-I{pybind11.get_include()}"]
    except Exception:
        pyb = []

    jobs, objs = [], []
    for f in cu:
        src = os.path.join(CSRC, f)
        flags = ARCH + NVCC_FLAGS + inc
        obj = os.path.join(BUILD, f"{f}.{_hash(src, flags)}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append([nvcc] + flags + ["-c", src, "-o", obj])
    for f in cpp:
        src = os.path.join(CSRC, f)
        needs_torch = "torch/" in open(src).read()
        flags = CXX_FLAGS + inc + pyb + [f"-I{sysconfig.get_paths()['include']}"] + (torch_inc + abi if needs_torch else [])
        obj = os.path.join(BUILD, f"{f}.{_hash(src, flags)}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append(["g++"] + flags + ["-c", src, "-o", obj])
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))

    out = so_path()
    stale = force or jobs or not os.path.exists(out) or any(
        os.path.getmtime(o) > os.path.getmtime(out) for o in objs)
    if stale:
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        link = ["g++", "-shared", "-o", out] + objs + [
            f"-L{tlib}", f"-L{os.path.join(cuda, 'lib64')}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
            "-ltorch", "-ltorch_python", "-lcudart", f"-Wl,-rpath,{tlib}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"]
        _run(link, verbose)
    # prune stale objects
    keep = set(objs)
    for f in os.listdir(BUILD):
        p = os.path.join(BUILD, f)
        if p.endswith(".o") and p not in keep:
            os.remove(p)
    return out


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", path)
