"""ctypes binding of ``libtetrad_b200.so`` (C ABI in ``include/tetrad_b200.h``).

The shared library is built in-tree by ``build()`` (``nvcc`` for sm_100a) and there is no
fallback of any kind: if it is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent
LIB_PATH = _PKG / "libtetrad_b200.so"
CSRC = _PKG / "csrc"
SOURCES = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [_ROOT / "include" / "tetrad_b200.h"]
BUILD_DIR = _PKG / "build"

# sm_100a only (no PTX for other architectures, no multi-arch fat binary); -lineinfo so that ncu's source page maps to
# these files
NVCC_COMPILE = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-cudart", "shared",
]
NVCC_LINK = ["-shared", "-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
NVCC_FLAGS = NVCC_COMPILE + NVCC_LINK

TQ_OK, TQ_EINVAL, TQ_ELOCUS, TQ_ECUDA, TQ_ENOMEM = 0, -1, -2, -3, -4

_lib = None


class TetradB200Error(RuntimeError):
    """A libtetrad_b200 call failed (status code + tq_last_error text)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libtetrad_b200 error {code}: {message}")
        self.code = code


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found; cannot build libtetrad_b200.so")
    return exe


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > built for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, defines=(), out: Path | None = None) -> Path:
    """Compile the CUDA sources for sm_100a into ``tetrad_b200/libtetrad_b200.so``: one ``nvcc -c`` per translation
    unit (in parallel, objects under ``tetrad_b200/build/``; a unit is recompiled when it or any header is newer
    than its object), then one link.  ``defines`` (e.g. ``["TQ_MMA_TIMERS"]``) and ``out`` build a variant library for
    measurements."""
    from concurrent.futures import ThreadPoolExecutor
    target = Path(out) if out else LIB_PATH
    if not force and not defines and out is None and not needs_build():
        return LIB_PATH
    objdir = BUILD_DIR / ("default" if not defines else "_".join(d.replace("=", "-") for d in defines))
    objdir.mkdir(parents=True, exist_ok=True)
    newest_header = max(p.stat().st_mtime for p in HEADERS)
    nvcc = _nvcc()

    def compile_one(src: Path):
        obj = objdir / (src.name + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, newest_header):
            return obj, ""
        cmd = [nvcc, *NVCC_COMPILE, *[f"-D{d}" for d in defines], "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name} ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
        return obj, proc.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        print("\n".join(err for _, err in results if err))
    proc = subprocess.run([nvcc, *NVCC_LINK, "-o", str(target), *[str(o) for o, _ in results]],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    return target


def _declare(lib):
    c = ctypes
    vp, i64, u64, i32 = c.c_void_p, c.c_int64, c.c_uint64, c.c_int
    lib.tq_version.restype = i32
    lib.tq_version.argtypes = []
    lib.tq_last_error.restype = c.c_char_p
    lib.tq_last_error.argtypes = []
    lib.tq_launch_count.restype = i64
    lib.tq_launch_count.argtypes = []
    lib.tq_last_kernel.restype = c.c_char_p
    lib.tq_last_kernel.argtypes = []
    lib.tq_last_count_warps.restype = i32
    lib.tq_last_count_warps.argtypes = []
    lib.tq_pairgemm_stats.restype = i32
    lib.tq_pairgemm_stats.argtypes = [c.POINTER(i64)]
    lib.tq_set_option.restype = i32
    lib.tq_set_option.argtypes = [c.c_char_p, i64]
    lib.tq_plan_create.restype = i32
    lib.tq_plan_create.argtypes = [vp, i64, i64, i32, vp, c.POINTER(vp)]
    lib.tq_plan_destroy.restype = None
    lib.tq_plan_destroy.argtypes = [vp]
    lib.tq_plan_info.restype = i32
    lib.tq_plan_info.argtypes = [vp, c.POINTER(i64)]
    lib.tq_plan_tables.restype = i32
    lib.tq_plan_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.tq_plan_create_runs.restype = i32
    lib.tq_plan_create_runs.argtypes = [vp, i64, i32, vp, c.POINTER(vp)]
    lib.tq_format_tsv.restype = i64
    lib.tq_format_tsv.argtypes = [vp, vp, vp, i64, vp, i64]
    lib.tq_round_scores.restype = i32
    lib.tq_round_scores.argtypes = [vp, i64, vp]
    lib.tq_format_qmc.restype = i64
    lib.tq_format_qmc.argtypes = [vp, vp, i64, vp, i64]
    lib.tq_format_qmc_rows.restype = i64
    lib.tq_format_qmc_rows.argtypes = [vp, vp, vp, i64, vp, i64]
    lib.tq_qmc_weights.restype = i64
    lib.tq_qmc_weights.argtypes = [vp, vp, vp, i64, i32, i64, c.c_double, vp, vp]
    lib.tq_unrank.restype = i32
    lib.tq_unrank.argtypes = [i32, vp, i64, i64, i64, vp, vp]
    lib.tq_sample_quartets.restype = i32
    lib.tq_sample_quartets.argtypes = [i32, i64, i64, i64, u64, vp, vp]
    lib.tq_sample_ranks_host.restype = i32
    lib.tq_sample_ranks_host.argtypes = [i64, i64, i64, u64, vp]
    lib.tq_boot_pack.restype = i32
    lib.tq_boot_pack.argtypes = [vp, vp, i64, i64, vp, c.c_uint32, c.c_uint32, vp, vp, vp]
    lib.tq_pack.restype = i32
    lib.tq_pack.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.tq_pack_rows.restype = i32
    lib.tq_pack_rows.argtypes = [vp, vp, i64, i64, i64, i64, vp, vp]
    lib.tq_boot_pack_rows.restype = i32
    lib.tq_boot_pack_rows.argtypes = [vp, vp, i64, i64, i64, i64, vp, c.c_uint32, c.c_uint32, vp, vp, vp]
    lib.tq_resolve.restype = i32
    lib.tq_resolve.argtypes = [vp, vp, i64, vp, i64, i32, u64, vp, vp, vp, vp, vp, vp]
    lib.tq_ctx_create.restype = i32
    lib.tq_ctx_create.argtypes = [i32, c.POINTER(vp)]
    lib.tq_ctx_destroy.restype = None
    lib.tq_ctx_destroy.argtypes = [vp]
    lib.tq_resolve_host.restype = i32
    lib.tq_resolve_host.argtypes = [vp, vp, i64, i64, i64, vp, i64, vp, i64, i32, u64, u64, vp, vp, vp, vp]
    lib.tq_ctx_last_traffic.restype = i32
    lib.tq_ctx_last_traffic.argtypes = [vp, c.POINTER(i64)]
    lib.tq_microbench.restype = i32
    lib.tq_microbench.argtypes = [i32, i32, c.POINTER(c.c_double)]


EXPORTS = [
    "tq_version", "tq_last_error", "tq_launch_count", "tq_last_kernel", "tq_last_count_warps", "tq_set_option", "tq_pairgemm_stats", "tq_plan_create", "tq_plan_create_runs", "tq_boot_pack", "tq_unrank", "tq_sample_quartets", "tq_sample_ranks_host",
    "tq_plan_destroy",
    "tq_plan_info", "tq_plan_tables", "tq_pack", "tq_pack_rows", "tq_boot_pack_rows", "tq_resolve", "tq_ctx_create", "tq_ctx_destroy", "tq_resolve_host",
    "tq_ctx_last_traffic", "tq_microbench", "tq_format_tsv", "tq_round_scores", "tq_format_qmc", "tq_format_qmc_rows", "tq_qmc_weights",
]


def load():
    """Load (never build) the shared library; raise if it is not there."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("TETRAD_B200_LIB", LIB_PATH))      # a variant build (measurements only)
        if not path.exists():
            raise RuntimeError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the CUDA extension is the only implementation; there is no CPU fallback)")
        lib = ctypes.CDLL(str(path), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
        _declare(lib)
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != TQ_OK:
        raise TetradB200Error(code, load().tq_last_error().decode("utf-8", "replace"))


KERNELS = {"auto": 0, "fused": 1, "mma": 3, "pairgemm": 5}


def set_kernel(name: str) -> None:
    """Force a resolve-kernel variant ("fused", "mma" = tensor-core count kernel, "pairgemm" = tip-pair GEMM on tcgen05
    for full mode) or go back to the size rule ("auto")."""
    check(load().tq_set_option(b"kernel", KERNELS[name]))


def pairgemm_stats() -> tuple[int, int]:
    """(chunks routed to the tip-pair GEMM, chunks routed to the per-quartet kernel) by full-mode launches since load;
    synchronise the device first."""
    out = (ctypes.c_int64 * 2)()
    check(load().tq_pairgemm_stats(out))
    return int(out[0]), int(out[1])
