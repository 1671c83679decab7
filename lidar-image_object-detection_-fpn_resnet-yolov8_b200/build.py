"""Builds libsfa_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and `python -m`.

    python "lidar-image_object-detection_-fpn_resnet-yolov8_b200/build.py" [--force] [--verbose]

No fast-math, and FMA contraction is off (`-fmad=false`): the parity-critical arithmetic (fp32
divide / floor / add) must round exactly like numpy does.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libsfa_b200.so")
SOURCES = ["bev_rasterize.cu", "bev_fused.cu", "peak_decode.cu", "box_projection.cu", "point_transform.cu", "lidar_filter.cu", "host_pipeline.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-prec-div=true",
    "-prec-sqrt=true", "-ftz=false",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


STAMP = LIB + ".flags"   # the variant (release / SFA_DEBUG_TIMING) the library on disk was built as


def _variant():
    # SFA_NVCC_DEFS: extra -D flags for kernel experiments (e.g. "-DSFA_FUSED_WORKERS=224"); part of the stamp
    base = "debug-timing" if os.environ.get("SFA_DEBUG_TIMING") == "1" else "release"
    extra = os.environ.get("SFA_NVCC_DEFS", "").strip()
    return base + (" " + extra if extra else "")


def _stale():
    if not os.path.exists(LIB):
        return True
    try:
        with open(STAMP) as f:
            if f.read().strip() != _variant():
                return True
    except OSError:
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "sfa_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = find_nvcc()
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-shared", "-o", LIB]
    if os.environ.get("SFA_DEBUG_TIMING") == "1":   # developer aid: per-phase cycle counters in bev_band
        cmd += ["-DSFA_DEBUG_TIMING"]
    cmd += os.environ.get("SFA_NVCC_DEFS", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-cudart=shared"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsfa_b200.so")
    with open(STAMP, "w") as f:
        f.write(_variant() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
