"""Build recipe for the CUDA extension (in-tree, sm_100a only).

    python -m brickmap_b200.build            # -> brickmap_b200/libbrickmap_b200.so

The flags matter for parity, not only for speed: -fmad=false stops nvcc from contracting mul+add pairs on its
own; every FMA the reference build performs is written out explicitly in csrc/bm_device.cuh.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["bm_kernels.cu", "bm_scene_store.cu"]
HEADERS = ["bm_device.cuh", "bm_frame_quantum.cuh", os.path.join("..", "..", "include", "brickmap_b200.h")]
LIB = os.path.join(HERE, "libbrickmap_b200.so")
NVCC_FLAGS = (["-DBM_QDEBUG"] if os.environ.get("BM_QDEBUG") else []) + ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-Xcompiler", "-ffp-contract=off",
              "-Xcompiler", "-fPIC", "-shared"]


def source_hash():
    """sha256 (first 12 hex digits) over the kernel sources: stamps measurements that describe one build (profiles/frame_kernel_traffic.json)."""
    import hashlib
    h = hashlib.sha256()
    for name in sorted(n for n in SOURCES + HEADERS if not n.startswith("..")):  # csrc/ only: the public header's comments do not change a kernel
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:12]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libbrickmap_b200.so with nvcc (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


def build_variant(name, defines, verbose=False):
    """A/B builds for measurements (tools/gpu_ab.sh, BRICKMAP_B200_LIB): libbrickmap_b200_<name>.so with extra -D switches."""
    out = os.path.join(HERE, "libbrickmap_b200_%s.so" % name)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m brickmap_b200.build --variant bulk BM_BULK_PROLOGUE=1 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if not a.startswith("-")], verbose="-v" in sys.argv))
    else:
        build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(LIB)
