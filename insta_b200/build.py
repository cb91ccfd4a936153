"""Builds libinsta_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libinsta_b200.so")
SOURCES = ["api.cu", "deform.cu", "hashgrid.cu", "nerf.cu", "mlp.cu", "mlp_bwd.cu", "train.cu", "trainer.cu", "occupancy.cu", "microbench.cu", "recorder.cu", "raycast.cu", "mesh_host.cpp", "snapshot.cpp"]
HEADERS = ["ctx.h", "grid_device.cuh", "nerf_device.cuh", "umma.cuh", os.path.join("..", "..", "include", "insta_b200.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _host_cxx():
    # the image's /opt/gcc wrapper lacks libgomp specs; the system g++ is the safe host compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    common = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-pthread", "--expt-relaxed-constexpr"]
    cxx = _host_cxx()
    if cxx:
        common += ["-ccbin", cxx]
    if verbose:
        common += ["-Xptxas", "-v"]
    if os.environ.get("INSTA_TUNE"):  # extra kernel variants for A/B measurements (tools/tune_deform.py)
        common += ["-DINSTA_TUNE"]
    if os.environ.get("INSTA_WQ_SLOTS"):
        common += ["-DINSTA_WQ_SLOTS=" + os.environ["INSTA_WQ_SLOTS"]]
    procs = []
    for s in srcs:
        o = os.path.join(LIB_DIR, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [_nvcc()] + common + ["-x", "cu", "-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-Xcompiler", "-pthread"] + (["-ccbin", cxx] if cxx else [])
    subprocess.check_call(cmd + ["-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
