"""Build libfluidmpm.so in-tree with nvcc for sm_100a (B200).  No torch headers are involved: the library
is a plain C-ABI shared object (include/fluidmpm.h, include/fluidsmoke.h) driven through ctypes."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = ["fmpm_forward.cu", "fmpm_backward.cu", "fmpm_io.cu", "fmpm_rigid.cu", "fsmk_smoke.cu"]
HDRS = ["fmpm_common.cuh", "fmpm_scatter.cuh", "fmpm_sdf.cuh", os.path.join("..", "..", "include", "fluidmpm.h"),
        os.path.join("..", "..", "include", "fluidsmoke.h")]
OUT = os.environ.get("FMPM_OUT", os.path.join(HERE, "..", "libfluidmpm.so"))   # FMPM_OUT: A/B variants (profiles/ab_variants.sh)
OBJDIR = os.environ.get("FMPM_OBJDIR", HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# no --use_fast_math: parity with the reference's IEEE fp32 arithmetic matters more than a few SFU cycles
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-DFMPM_BUILD",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "--cudart", "static"] + os.environ.get("FMPM_DEFS", "").split()


def needs_build():
    out = os.path.abspath(OUT)
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SRCS + HDRS + ["build.py"])


def build(force=False, verbose=False):
    out = os.path.abspath(OUT)
    if not force and not needs_build():
        return out
    objs = []
    procs = []
    os.makedirs(OBJDIR, exist_ok=True)
    for s in SRCS:
        o = os.path.join(OBJDIR, s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    subprocess.check_call([NVCC, "-shared", "-o", out, "--cudart", "static", "-ccbin", "/usr/bin/g++"] + objs)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
