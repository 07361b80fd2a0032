"""Build recipe for the REAL reference CUDA extension (test infrastructure only).

Compiles /root/reference/gs/src/{render.cu,bindings.cpp} -- where they lie, via a scratch
copy under /tmp because HEAD does not compile -- into ``oracle/_ref/_gs_ref*.so`` for sm_100a.
Nothing from /root/reference is copied into the repository; ``oracle/_ref/`` is git-ignored
(but travels to the GPU box with gpurun).

Scratch-copy patch (the only change; SURVEY.md "Facts"): ``vol_render_bg.h`` assigns to a
local ``out[]`` before it is declared (2 kernels, 6 lines).  The forward kernel's three lines are
re-pointed at ``out_rgb`` (the evident intent: empty tile -> background colour); the backward
kernel's three lines are dropped (they would write a forward output inside backward).

Flags: ``-O3 -std=c++17`` (torch 2.11 headers need c++17; reference says c++14), ``-DNDEBUG``
(device asserts off: the reference's bitwise ``assert(out_rgb == out)`` between two separately
compiled kernels would abort the process; it also gives the baseline its best speed),
``-gencode arch=compute_100a,code=sm_100a``.

Usage:  python oracle/build_ref.py        (no-op when /root/reference is absent or .so is fresh)
"""
import os
import re
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path("/root/reference/gs/src")
NAME = "_gs_ref"


def existing():
    if not OUT.exists():
        return None
    for p in OUT.glob(NAME + "*.so"):
        return p
    return None


def build(verbose=False, force=False):
    so = existing()
    if so is not None and not force:
        return so
    if not REF.exists():
        return None
    OUT.mkdir(exist_ok=True)
    scratch = Path(tempfile.mkdtemp(prefix="gsref_"))
    src = scratch / "src"
    shutil.copytree(REF, src)
    bg = src / "include" / "vol_render_bg.h"
    text = bg.read_text().split("\n")
    seen = 0
    for i, line in enumerate(text):
        m = re.match(r"^(\s*)out\[(\d)\] = bg_rgb\[(\d)\];\s*$", line)
        if not m:
            continue
        # occurrences 0..2 = forward kernel, 3..5 = backward kernel (both precede `float out[3]`)
        if seen < 3:
            text[i] = (f"{m.group(1)}out_rgb[3 * (global_y * W + global_x) + {m.group(2)}]"
                       f" = bg_rgb[{m.group(3)}];")
        elif seen < 6:
            text[i] = ""
        seen += 1
        if seen == 6:
            break
    assert seen == 6, f"expected 6 pre-declaration uses of out[], found {seen}"
    bg.write_text("\n".join(text))

    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    import torch  # noqa: F401
    from torch.utils.cpp_extension import load

    build_dir = scratch / "build"
    build_dir.mkdir()
    load(
        name=NAME,
        sources=[str(src / "render.cu"), str(src / "bindings.cpp")],
        extra_include_paths=[str(src / "include")],
        extra_cflags=["-O3", "-std=c++17", "-DNDEBUG"],
        extra_cuda_cflags=["-O3", "-std=c++17", "-DNDEBUG",
                           "-gencode", "arch=compute_100a,code=sm_100a",
                           "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                           "-U__CUDA_NO_HALF2_OPERATORS__"],
        build_directory=str(build_dir),
        verbose=verbose,
        is_python_module=False,
    )
    built = list(build_dir.glob(NAME + "*.so"))
    assert built, "reference extension did not produce a .so"
    dst = OUT / built[0].name
    shutil.copy2(built[0], dst)
    shutil.rmtree(scratch, ignore_errors=True)
    return dst


REF_ROOT = Path("/root/reference")
PY_OUT = OUT / "py"


def stage_python(force=False):
    """Stage the reference's UNMODIFIED Python packages (gs/*.py, utils/**/*.py) into the git-ignored
    ``oracle/_ref/py`` so that tests on the GPU box can run the reference's own SHRenderer.forward /
    backward and its training loop over this repo's ``_gs`` shim (INTEGRATION.md route 2).  Byte-for-byte
    copies, never committed (``oracle/_ref/`` is in .gitignore), test infrastructure only."""
    if not REF_ROOT.exists():
        return PY_OUT if PY_OUT.exists() else None
    marker = PY_OUT / ".staged"
    if marker.exists() and not force:
        return PY_OUT
    if PY_OUT.exists():
        shutil.rmtree(PY_OUT)
    for pkg in ("gs", "utils"):
        src_dir = REF_ROOT / pkg
        for src in src_dir.rglob("*.py"):
            rel = src.relative_to(REF_ROOT)
            if rel.parts[:2] == ("gs", "src"):
                continue
            dst = PY_OUT / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copy2(src, dst)
    marker.write_text("staged from /root/reference (unmodified)\n")
    return PY_OUT


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("reference extension:", p)
    print("reference python:", stage_python(force="-f" in sys.argv))
