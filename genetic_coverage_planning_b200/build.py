"""Build libgcp.so (sm_100a) in-tree with nvcc.  `python -m genetic_coverage_planning_b200.build`

Two flavours:
  libgcp.so        release (what the product loads)
  libgcp_debug.so  the same sources with -DGCP_BOUNDS_CHECK: every shared-memory / scratch index the
                   kernels form is checked with a device-side assert before use.  compute-sanitizer is
                   not available on the GPU pool, so this build is what tests/test_gpu_debug_build.py
                   runs the whole kernel family under (select it with GCP_LIB=debug).
Sources are compiled to objects in parallel and linked; objects are reused while their inputs are
unchanged."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgcp.so")
LIB_DEBUG = os.path.join(HERE, "libgcp_debug.so")
OBJDIR = os.path.join(HERE, "_build")
SOURCES = ["gcp_api.cu", "gcp_eval.cu", "gcp_frustum.cu", "gcp_graph.cu", "gcp_fused.cu", "gcp_host.cu",
           "gcp_lcbig.cu", "gcp_rank.cu", "gcp_sort.cu", "gcp_vary.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(os.path.dirname(HERE), "include", "gcp.h")]


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in _deps() + [os.path.join(CSRC, s) for s in SOURCES])


def _compile(src, obj, extra, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False, debug=False):
    lib = LIB_DEBUG if debug else LIB
    if not force and not needs_build(lib):
        return lib
    os.makedirs(OBJDIR, exist_ok=True)
    extra = ["-DGCP_BOUNDS_CHECK"] if debug else ["-DNDEBUG"]
    tag = "dbg" if debug else "rel"
    hdr_t = max(os.path.getmtime(d) for d in _deps())
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, "%s.%s.o" % (s[:-3], tag))
        objs.append(obj)
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(src)):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1) or 1) as ex:
        for src, rc, out in ex.map(lambda j: _compile(j[0], j[1], extra, verbose), jobs):
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError("nvcc failed on %s" % src)
            if verbose:
                sys.stderr.write(out)
    r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking %s" % lib)
    return lib


def build_variant(tag, defines, sources=("gcp_fused.cu",)):
    """libgcp_<tag>.so = the release objects with `sources` recompiled under extra -D switches
    (A/B runs of kernel experiments: GCP_LIB=<tag> python profiles/exp_fused.py ...)."""
    build()
    objs = []
    for s in SOURCES:
        obj = os.path.join(OBJDIR, "%s.rel.o" % s[:-3])
        if s in sources:
            obj = os.path.join(OBJDIR, "%s.%s.o" % (s[:-3], tag))
            src, rc, out = _compile(os.path.join(CSRC, s), obj, ["-DNDEBUG"] + ["-D" + d for d in defines], False)
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    lib = os.path.join(HERE, "libgcp_%s.so" % tag)
    r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking %s" % lib)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:          # --variant tag KF_X_A=0 KF_X_B=1 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
