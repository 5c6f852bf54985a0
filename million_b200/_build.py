"""Build libmillion_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python -m million_b200._build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmillion_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("MILLION_NVCC_EXTRA", "").split()      # e.g. MILLION_NVCC_EXTRA=-DMILLION_FUSED_SPLITKV (experimental option)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "million_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())     # names, not absolute paths: the stamp must hold on the GPU box's copy of the tree
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, out: str = None) -> str:
    """`out`: build a VARIANT (MILLION_NVCC_EXTRA switches) into that path, with its own object directory, leaving the in-tree
    library and its stamp alone (A/B runs: MILLION_B200_LIB=variants/x.so)."""
    dig = _digest()
    variant = out is not None
    out = out or OUT
    if not variant and not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objs = []
    build_dir = os.path.join(HERE, "build") if not variant else out + ".build"
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src) + ".o")
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        text, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{text}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcuda"]
    subprocess.check_call(link)
    with open(os.path.join(build_dir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if not variant:
        with open(STAMP, "w") as f:
            f.write(dig)
    if verbose:
        print("\n".join(log))
    return out


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out))
