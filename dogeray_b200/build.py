"""Build libdogeray_b200.so in-tree with nvcc for sm_100a.

    python dogeray_b200/build.py [--force] [--verbose]

Flags that matter (DESIGN.md "Arithmetic"):
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX fallback, no multi-arch
  -fmad=false                               single IEEE operations, so integer outputs (Morton keys,
                                            hit ids) reproduce on the host bit for bit; the box test
                                            requests its FMAs explicitly with fmaf()
  -lineinfo                                 ncu source page
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdogeray_b200.so")
STAMP = os.path.join(HERE, ".libdogeray_b200.stamp")

SOURCES = ["rts_loader.cpp", "image_io.cpp", "scene.cu", "render.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall,-Wno-unused-function,-pthread",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
] + os.environ.get("DRB_NVCC_EXTRA", "").split()      # e.g. -DDRB_TRACE_STEPS=3 for tuning experiments


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    names = sorted(os.listdir(CSRC)) + ["../../include/dogeray_b200.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed; returns the path of the shared library."""
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print("+", " ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed building libdogeray_b200.so")
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"]
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    # headless CLI, the drop-in for `raygpu.exe scene.rts`
    cli = os.path.join(HERE, "dogeray-b200")
    cmd = ["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(CSRC, "cli.cpp"), "-o", cli,
           "-L" + HERE, "-ldogeray_b200", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(path)
