"""Build ``libse_b200.so`` in-tree with nvcc for sm_100a (no torch / libtorch linkage).

``python -m speech_enhancement_by_s3prl_b200.build`` or ``build_library()``.  The shared
object is written next to this file so that it travels with the source tree; it is
rebuilt when any file under ``csrc/`` or ``include/`` is newer than the library.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libse_b200.so")
HASH_PATH = LIB_PATH + ".sha256"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libse_b200.so cannot be built (there is no CPU fallback)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash():
    """sha256 over csrc/ and include/ (file names + contents): mtimes do not survive the copy to the GPU box."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(ROOT, "include", "*.h"))):
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())
            with open(path, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build_library(force=False, verbose=False, extra_flags=(), out_path=None):
    """extra_flags / out_path: experiment builds (e.g. -DSE_K3_MIN_BLOCKS=4 into another file, loaded via SE_B200_LIB)."""
    if out_path is None and not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        if out_path is not None:
            obj = obj[:-2] + ".exp.o"
        cmd = [nvcc, *ARCH_FLAGS, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", *extra_flags,
               "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    target = out_path or LIB_PATH
    tmp = target + ".tmp"
    subprocess.check_call([nvcc, *ARCH_FLAGS, "-shared", "-o", tmp, *objs, "-lcudart"])
    os.replace(tmp, target)
    if out_path is None:
        with open(HASH_PATH, "w") as f:
            f.write(source_hash())
    return target


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
