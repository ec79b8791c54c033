"""Build libagatha_b200.so (C ABI, include/agatha_b200.h) in-tree with nvcc for sm_100a.

    python -m agatha_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libagatha_b200.so")
STAMP = os.path.join(LIBDIR, "libagatha_b200.stamp")
SYNTH_LIB = os.path.join(LIBDIR, "libagatha_synth.so")

CUDA_SOURCES = ["engine.cu", "stream.cu", "int_peak.cu", "extend_inst_c2_c4.cu", "extend_inst_c8_c16.cu", "extend_inst_c24.cu", "extend_inst_c32.cu",
                "extend_inst_wide2.cu", "extend_inst_wide4.cu", "extend_inst_wide8.cu",
                "extend16_inst_c8_c16.cu", "extend16_inst_c24.cu", "extend16_inst_c32.cu",
                "extend16_inst_wide2.cu", "extend16_inst_wide4.cu", "extend16_inst_wide8.cu"]
CXX_SOURCES = ["host_utils.cpp", "host_pack.cpp", "fasta.cpp", "job.cpp", "gasal_compat.cpp"]   # manual_main.cpp is the driver, linked separately
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp,-Wall", "--default-stream", "per-thread"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _sources():
    return [os.path.join(CSRC, f) for f in CUDA_SOURCES + CXX_SOURCES if os.path.exists(os.path.join(CSRC, f))]


def _digest():
    h = hashlib.sha256()
    files = sorted([os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h"))])
    files += [os.path.join(ROOT, "tools", "synth", f) for f in ("synth.cpp", "agatha_synth.h")]
    for dp, _, fns in sorted(os.walk(os.path.join(ROOT, "include"))):
        files += [os.path.join(dp, f) for f in sorted(fns)]
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(SYNTH_LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src) + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-c", src, "-o", obj]
        if src.endswith(".cpp"):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-x", "cu", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        if verbose and out.strip():
            print(out)
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    # the command-line driver (reference contract: test_prog.cpp / args_parser.cpp), linked against the library
    bindir = os.path.join(PKG, "bin")
    os.makedirs(bindir, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(CSRC, "manual_main.cpp"), "-o", os.path.join(bindir, "agatha_manual"),
           "-L" + LIBDIR, "-lagatha_b200", "-Wl,-rpath,$ORIGIN/../lib"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("driver link failed:\n" + r.stdout)
    # bench / test tooling: the synthetic workload generator, deliberately NOT part of the product library
    synth = os.path.join(ROOT, "tools", "synth")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fopenmp", "-I" + synth, os.path.join(synth, "synth.cpp"), "-o", SYNTH_LIB]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("synth build failed:\n" + r.stdout)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
