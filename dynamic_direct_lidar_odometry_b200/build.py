"""Build libddlo_gicp_b200.so (hand-written sm_100a CUDA + the extern "C" ABI) in-tree with nvcc.

Every .cu is compiled to its own object (in parallel, objects cached under _build/ and reused while the source, the
headers and the flags are unchanged), then linked into the shared library.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
OBJ_DIR = PKG / "_build"
LIB_PATH = LIB_DIR / "libddlo_gicp_b200.so"
SOURCES = ["api.cu", "prims.cu", "index.cu", "knn_cov.cu", "gicp.cu", "preprocess.cu", "cluster_sort.cu", "segmentation.cu", "batch.cu", "batch_align.cu",
           "keyframes.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _headers():
    return sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h"))


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*.cu")) + _headers())
    return LIB_PATH.stat().st_mtime < newest


def _stamp(src: Path, flags) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for p in [src] + _headers():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build_variant(name: str, extra_flags) -> Path:
    """a tuning variant of the library (extra -D flags) as lib/variants/<name>.so; select it with DDLO_GICP_LIB"""
    out_dir = LIB_DIR / "variants"
    out_dir.mkdir(parents=True, exist_ok=True)
    obj_dir = OBJ_DIR / ("variant_" + name)
    obj_dir.mkdir(parents=True, exist_ok=True)
    nvcc = nvcc_path()
    flags = [*NVCC_FLAGS, *extra_flags]

    def one(src_name):
        obj = obj_dir / (Path(src_name).stem + ".o")
        res = subprocess.run([nvcc, *flags, "-c", "-o", str(obj), str(CSRC / src_name)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(res.stdout + res.stderr)
        return str(obj)

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(one, SOURCES))
    out = out_dir / f"{name}.so"
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stdout + res.stderr)
    return out


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """force=True relinks and recompiles every object whose inputs changed (objects are content-addressed, so an
    unchanged source is never recompiled needlessly); DDLO_REBUILD_ALL=1 ignores the object cache."""
    if not force and not needs_build():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    extra = os.environ.get("DDLO_NVCC_EXTRA", "").split()  # e.g. -DDDLO_VISIT_STATS for profiles/visit_stats.py
    flags = [*NVCC_FLAGS, *extra] + (["-Xptxas=-v"] if verbose else [])
    nvcc = nvcc_path()
    rebuild_all = os.environ.get("DDLO_REBUILD_ALL") == "1" or verbose

    def compile_one(name: str):
        src = CSRC / name
        obj = OBJ_DIR / (src.stem + ".o")
        tag = OBJ_DIR / (src.stem + ".stamp")
        stamp = _stamp(src, flags)
        if not rebuild_all and obj.exists() and tag.exists() and tag.read_text() == stamp:
            return name, 0, ""
        res = subprocess.run([nvcc, *flags, "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
        if res.returncode == 0:
            tag.write_text(stamp)
        return name, res.returncode, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    failed = [(n, out) for n, rc, out in results if rc != 0]
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"--- {n}\n{out}" for n, out in failed))
    if verbose:
        for n, _, out in results:
            print(f"--- {n}\n{out}")
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
                          *[str(OBJ_DIR / (Path(s).stem + ".o")) for s in SOURCES]], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_library(force=True, verbose="-v" in sys.argv))
