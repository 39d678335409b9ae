"""Build the REFERENCE's own Mamba-1 CUDA extension (CrossMamba/FusionMamba/selective_scan/*.cu, unmodified, compiled where the
sources lie) for sm_100a into git-ignored baseline/_ref/, so that tools/microbench_sscan.py can put the "kernel to beat"
(reference selective_scan_fwd_kernel.cuh / selective_scan_bwd_kernel.cuh: one CTA per (batch, channel) row, CUB block scans) next to
libb200ssm's numbers on the same B200.  Needs /root/reference (build container); the resulting .so travels to the GPU box.
    python tools/build_ref_cuda.py
"""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("REFERENCE_ROOT", "/root/reference") + "/CrossMamba/FusionMamba/selective_scan"
OUT = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(SRC):
        print("reference sources not present; nothing built")
        return 1
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load
    srcs = [os.path.join(SRC, "selective_scan.cpp")] + sorted(glob.glob(os.path.join(SRC, "*.cu")))
    mod = load(name="selective_scan_cuda", sources=srcs, build_directory=OUT, verbose=True, with_cuda=True,
               extra_cflags=["-O3", "-std=c++17"],
               extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                                  "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                                  "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                                  "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math", "--ptxas-options=-v", "-lineinfo",
                                  "-gencode", "arch=compute_100a,code=sm_100a"],
               is_python_module=False)
    print("built", mod)
    return 0


if __name__ == "__main__":
    sys.exit(main())
