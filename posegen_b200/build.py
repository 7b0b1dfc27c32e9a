"""Build libposegen_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m posegen_b200.build [--verbose]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libposegen_b200.so")
SOURCES = ["pgn_api.cu", "pgn_stage_kernels.cu", "pgn_render_fp32.cu", "pgn_render_bf16.cu", "pgn_render_bf16_fc.cu", "pgn_train_kernels.cu", "pgn_delta_chain.cu", "pgn_batch_kernels.cu", "pgn_wgrad.cu", "pgn_input_grads.cu", "pgn_probe.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "posegen_b200.h"))
    return deps


def build(verbose: bool = False, force: bool = False, extra_flags=(), out: str | None = None) -> str:
    """extra_flags / out: experiment builds (`-DPGN_EXP_...`) into another file, loaded with POSEGEN_B200_LIB=<path>."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if out is not None:
        return _build_variant(list(extra_flags), out, verbose)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest(sources()):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building posegen_b200 (see stderr)")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    subprocess.run(link, check=True)
    return LIB_PATH


def _build_variant(flags, out, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = os.path.splitext(os.path.basename(out))[0]
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, f"{tag}_{src.replace('.cu', '.o')}")
        procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, *flags, "-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{o}")
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs], check=True)
    for o in objs:
        os.remove(o)
    return out


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv, force="--force" in sys.argv))
