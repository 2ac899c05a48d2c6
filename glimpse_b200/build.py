"""Build ``libglimpse_b200.so`` in-tree with nvcc for sm_100a (``python -m glimpse_b200.build``)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "glimpse_b200.cu")
OUT = os.path.join(HERE, "libglimpse_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("glimpse_b200.cu", "common.cuh", "camera.cuh", "motion.cuh", "tile.cuh",
                                                 "median25.cuh", "stream.cuh", "viewshed.cuh")] + [os.path.join(HERE, "..", "include", "glimpse_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "--use_fast_math=false" if False else "-Xptxas=-v"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    built = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > built for d in DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    """``defines`` / ``out``: a tuning variant (``-DNAME=VALUE`` flags) written next to the library under another name."""
    if not force and out == OUT and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libglimpse_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="--quiet" not in sys.argv, defines=defs,
                out=os.path.join(HERE, outs[0]) if outs else OUT))
