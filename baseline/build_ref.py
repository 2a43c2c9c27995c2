#!/usr/bin/env python
"""Build the UNMODIFIED reference CUDA extensions for sm_100a into baseline/_ref/ (git-ignored).

    python baseline/build_ref.py            # needs /root/reference (this container only)

The sources are compiled where they lie under /root/reference (nothing is copied into the repo);
only the two extension modules land in baseline/_ref/:

    selective_scan_cuda.so   mamba/csrc/selective_scan/*.cu + selective_scan.cpp   (mamba/setup.py:123-161)
    causal_conv1d_cuda.so    causal-conv1d/csrc/*.cu + causal_conv1d.cpp           (causal-conv1d/setup.py)

Flags are the reference's own (mamba/setup.py:139-156) with the one change SURVEY.md 8(c) records:
`-gencode arch=compute_100a,code=sm_100a` instead of the hard-coded sm_70/80/90 list.  They are the
same-box GPU comparison point of bench.py (`ref_cuda_us`); no product path imports them.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VIVIM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "baseline", "_ref")

NVCC = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
        "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
        "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
        "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++"]


def build(force=False):
    if not os.path.isdir(REF):
        print(f"{REF} is absent: nothing to build (the GPU box uses the prebuilt files)")
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ["CC"], os.environ["CXX"] = "/usr/bin/gcc", "/usr/bin/g++"
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    if not force and all(os.path.exists(os.path.join(OUT, n + ".so")) for n in ("selective_scan_cuda", "causal_conv1d_cuda")):
        return True
    scan_dir = os.path.join(REF, "mamba", "csrc", "selective_scan")
    conv_dir = os.path.join(REF, "causal-conv1d", "csrc")
    jobs = (
        ("selective_scan_cuda", scan_dir,
         ["selective_scan.cpp"] + [f"selective_scan_{k}.cu" for k in (
             "fwd_fp32", "fwd_fp16", "fwd_bf16", "bwd_fp32_real", "bwd_fp32_complex", "bwd_fp16_real",
             "bwd_fp16_complex", "bwd_bf16_real", "bwd_bf16_complex")]),
        ("causal_conv1d_cuda", conv_dir,
         ["causal_conv1d.cpp", "causal_conv1d_fwd.cu", "causal_conv1d_bwd.cu", "causal_conv1d_update.cu"]),
    )
    for name, src_dir, files in jobs:
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=[os.path.join(src_dir, f) for f in files], extra_include_paths=[src_dir],
             extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=NVCC, build_directory=bdir,
             is_python_module=False, verbose=True)
        so = os.path.join(OUT, name + ".so")
        os.replace(os.path.join(bdir, name + ".so"), so)
        shutil.rmtree(bdir, ignore_errors=True)          # objects are not needed on the GPU box
        subprocess.run(["strip", "--strip-debug", so], check=False)
        print("built", so)
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)
