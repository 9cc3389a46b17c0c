"""Build libvitb200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python vit-cifar_b200/build.py [--force]

The library is compiled for sm_100a only (`-gencode arch=compute_100a,code=sm_100a -lineinfo`); nvcc
cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the repo
snapshot, so nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.environ.get("VITB_BUILD_OUT") or os.path.join(HERE, "libvitb200.so")

SOURCES = ["elementwise.cu", "loss_adam.cu", "gemm_simt.cu", "gemm_tc.cu", "attention.cu", "head.cu", "dp.cu"]
HEADERS = ["common.cuh", "gemm_internal.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libvitb200.so cannot be built")
    return cand


def _newest_input() -> float:
    paths = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(INCLUDE, "vitb200.h"), __file__]
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _newest_input()


EXTRA_DEFINES = [d for d in os.environ.get("VITB_BUILD_DEFINES", "").split() if d]  # e.g. "-DVITB_PDL_EARLY_CTAS=0" (experiments)


def _compile(src: str) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    cmd = [_nvcc(), *NVCC_FLAGS, *EXTRA_DEFINES, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ_DIR, src.replace(".cu", ".ptxas.log"))
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    tmp = LIB_PATH + ".tmp"
    cmd = [_nvcc(), "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    try:
        write_sass_summary()
    except Exception as ex:  # evidence file only: never fail the build over it
        print(f"sass summary not written: {ex}")
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


SASS_SUMMARY = os.path.join(os.path.dirname(HERE), "profiles", "sass_summary.txt")
_MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDSM", "STSM", "MUFU.EX2", "REDG", "ATOMG"]


def write_sass_summary(path: str = SASS_SUMMARY) -> str:
    """Per-kernel counts of the SASS mnemonics that show what a kernel is made of (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
    UTMALDG / UTMASTG / UTMAPF = TMA load / store / prefetch, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS =
    mbarrier, HMMA = mma.sync, LDSM / STSM = ldmatrix / stmatrix.  Regenerated after every build of the library."""
    import re
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    r = subprocess.run([cuobjdump, "-sass", LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        return ""
    demangle = shutil.which("c++filt")
    rows, name, counts, n_instr = [], None, None, 0
    for line in r.stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name is not None:
                rows.append((name, n_instr, counts))
            name, counts, n_instr = m.group(1), {k: 0 for k in _MNEMONICS}, 0
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m:
            n_instr += 1
            ins = m.group(1)
            for k in _MNEMONICS:
                if re.search(r"(^|\s)" + re.escape(k) + r"(\.|\s|$)", ins):
                    counts[k] += 1
    if name is not None:
        rows.append((name, n_instr, counts))
    if demangle:
        d = subprocess.run([demangle], input="\n".join(n for n, _, _ in rows), capture_output=True, text=True).stdout.splitlines()
        if len(d) == len(rows):
            rows = [(dn, ni, c) for dn, (_, ni, c) in zip(d, rows)]
    rows.sort(key=lambda t: t[0])
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("SASS mnemonic counts per kernel of vit-cifar_b200/libvitb200.so (sm_100a), written by vit-cifar_b200/build.py\n")
        f.write("UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG/UTMAPF = TMA load/store/prefetch, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,\n")
        f.write("SYNCS = mbarrier ops, HMMA = mma.sync, LDSM/STSM = ldmatrix/stmatrix; only non-zero counts are listed\n\n")
        tot = {k: 0 for k in _MNEMONICS}
        for nm, ni, c in rows:
            short = re.sub(r"\(.*$", "", nm)
            nz = "  ".join(f"{k}={v}" for k, v in c.items() if v)
            f.write(f"{short}\n    instructions={ni}  {nz}\n")
            for k, v in c.items():
                tot[k] += v
        f.write("\nTOTAL  " + "  ".join(f"{k}={v}" for k, v in tot.items()) + f"  kernels={len(rows)}\n")
    return path


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
