"""Build the CUDA library in-tree (robustgrape_b200/lib/librobustgrape_b200.so) for sm_100a."""
from __future__ import annotations

import os
import re
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "lib" / "librobustgrape_b200.so"
COMPILE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                 "-diag-suppress", "68,128,20058", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
OBJ_DIR = ROOT / "lib" / "obj"


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(CSRC.glob("*.cu"))


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(src, seen=None):
    """The translation unit and every project header it includes (recursively)."""
    seen = set() if seen is None else seen
    src = src.resolve()
    if src in seen or not src.exists():
        return seen
    seen.add(src)
    for inc in _INC.findall(src.read_text()):
        _deps((src.parent / inc), seen)
    return seen


def _stale(src, obj):
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _deps(src))


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for src in sources() for p in _deps(src))


def build_cuda(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a (one nvcc process per translation unit, in parallel) and link the
    in-tree shared library."""
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in sources():
        obj = OBJ_DIR / (src.stem + ".o")
        if not force and not _stale(src, obj):      # per-translation-unit incremental build
            objs.append(str(obj))
            continue
        cmd = [nvcc, *COMPILE_FLAGS, "-c", "-o", str(obj), str(src)]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            for _, _, q in procs:
                if q.poll() is None:
                    q.kill()
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
        objs.append(str(obj))
    r = subprocess.run([nvcc, *LINK_FLAGS, "-o", str(LIB), *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB
