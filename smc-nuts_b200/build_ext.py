"""Build libsmcnuts_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/smcnuts_b200.h) in-tree.

    python smc-nuts_b200/build_ext.py [--force] [-v]

nvcc cross-compiles without a GPU.  The shared objects land in smc-nuts_b200/smcnuts/_lib/ (git-ignored,
but they travel to the GPU box with the repo snapshot):

  libsmcnuts_b200.so          the product library
  libsmcnuts_b200_parity.so   the PARITY build of the same sources: -DSMCB_PARITY=1 -fmad=false, i.e. the oracle's
                              statement order in the model device functions and no FMA contraction.  Same C-ABI; never
                              loaded by the product path -- the `-m gpu` parity tests load it to show that the CUDA
                              binary reproduces the oracle tree for tree and bit for bit.
"""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT_DIR = HERE / "smcnuts" / "_lib"
OUT = OUT_DIR / "libsmcnuts_b200.so"
OUT_PARITY = OUT_DIR / "libsmcnuts_b200_parity.so"
PARITY_FLAGS = ["-DSMCB_PARITY=1", "-fmad=false"]
SOURCES = ["nuts_kernel.cu", "nuts_kernel_scaled.cu", "weights.cu", "resample.cu", "gauss_lkernel.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "--use_fast_math=false"]
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]  # IEEE FP64 everywhere; no fast-math


def _stale(target, deps):
    return (not target.exists()) or target.stat().st_mtime < max(p.stat().st_mtime for p in deps)


def build(force=False, verbose=False, parity=True):
    out = _build_one(OUT, HERE / "build", [], force, verbose)
    if parity:
        _build_one(OUT_PARITY, HERE / "build_parity", PARITY_FLAGS, force, False)
    return out


def _build_one(OUT, obj_dir, extra_flags, force, verbose):
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    obj_dir.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "smcnuts_b200.h"]

    def compile_one(src):
        obj = obj_dir / (src[:-3] + ".o")
        if force or _stale(obj, [CSRC / src] + headers):
            cmd = ["nvcc", *NVCC_FLAGS, *extra_flags, "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(OUT, objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(OUT), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


def build_variant(tag, flags):
    """A/B experiments: another build of the same sources with extra -D flags -> _lib/libsmcnuts_b200_<tag>.so
    (select it with SMCB_LIB_PATH; tools/ab_time.py)."""
    return _build_one(OUT_DIR / f"libsmcnuts_b200_{tag}.so", HERE / f"build_{tag}", list(flags), False, False)


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
