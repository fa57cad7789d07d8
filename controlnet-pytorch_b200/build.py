"""In-tree build of libcnb200.so (hand-written CUDA for sm_100a) and of the C-ABI export check.

    python controlnet-pytorch_b200/build.py [--force]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so it travels to the GPU box with the
gpurun snapshot (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libcnb200.so")
STAMP = os.path.join(HERE, ".libcnb200.stamp")
SOURCES = ["api.cu", "conv_f32.cu", "conv_small.cu", "conv_tma.cu", "conv_tma_f16.cu", "conv_tma_tf32.cu", "groupnorm.cu", "groupnorm_big.cu", "attention_f32.cu", "attention_f16.cu", "attention_tmem.cu", "elementwise.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-shared"]
# --use_fast_math only affects the plain operators; every bit-exactness-critical expression in elementwise.cu
# uses the explicit *_rn intrinsics, sinf/cosf/expf/logf calls there are kept accurate via -fmad only (see below).


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(INCLUDE, "cnb200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return LIB
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)

    def compile_one(src):
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        f = list(flags)
        if src == "elementwise.cu":
            f.remove("--use_fast_math")      # accurate sinf/cosf/logf/expf for the embedding + scheduler kernels
        cmd = [_nvcc()] + f + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    from concurrent.futures import ThreadPoolExecutor
    objs = []
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:   # one nvcc per source file
        for src, obj, r in pool.map(compile_one, SOURCES):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on " + src)
            objs.append(obj)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
