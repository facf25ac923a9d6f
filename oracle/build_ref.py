"""Build the reference's own native extensions (unmodified sources, compiled where
they lie under /root/reference) into oracle/_ref/ for use as a parity checker.

TEST INFRASTRUCTURE ONLY -- nothing under isopoints_b200/ may import this.

Outputs (git-ignored, but they travel to the GPU box with gpurun):
  oracle/_ref/ref_frnn_C/ref_frnn_C.so       <- external/FRNN/frnn/csrc/**   (frnn._C)
  oracle/_ref/ref_prefix_sum/ref_prefix_sum.so <- external/FRNN/external/prefix_sum
  oracle/_ref/ref_dss_C/ref_dss_C.so          <- DSS/csrc/{ext.cpp,rasterize_points*.cu,rasterize_points_cpu.cpp}

The only addition is an EMPTY shim header THC/THCNumerics.cuh (removed from modern
torch; rasterize_points.cu:5 includes it but uses nothing from it), placed in
oracle/_ref/shim/ -- no reference file is edited or copied.
"""
import os
import sys
import glob

REF = os.environ.get("ISO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(which=("frnn", "dss", "prefix_sum"), verbose=False):
    if not os.path.isdir(REF):
        print("reference tree not present; using prebuilt oracle/_ref if any")
        return
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 8))
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(OUT, "shim", "THC")
    os.makedirs(shim, exist_ok=True)
    open(os.path.join(shim, "THCNumerics.cuh"), "a").close()

    if "frnn" in which:
        c = os.path.join(REF, "external/FRNN/frnn/csrc")
        srcs = [os.path.join(c, "ext.cpp")]
        srcs += sorted(glob.glob(os.path.join(c, "*", "*.cpp")))
        srcs += [s for s in sorted(glob.glob(os.path.join(c, "*", "*.cu")))
                 if not s.endswith("grid/prefix_sum.cu")]  # unbound dead code (ext.cpp)
        d = os.path.join(OUT, "ref_frnn_C"); os.makedirs(d, exist_ok=True)
        load(name="ref_frnn_C", sources=srcs, extra_include_paths=[c],
             extra_cflags=["-O2"], extra_cuda_cflags=["-O2"],
             build_directory=d, with_cuda=True, verbose=verbose)
    if "dss" in which:
        c = os.path.join(REF, "DSS/csrc")
        srcs = [os.path.join(c, f) for f in
                ("ext.cpp", "rasterize_points.cu", "rasterize_points_backward.cu",
                 "rasterize_points_cpu.cpp")]
        d = os.path.join(OUT, "ref_dss_C"); os.makedirs(d, exist_ok=True)
        load(name="ref_dss_C", sources=srcs,
             extra_include_paths=[c, os.path.join(OUT, "shim")],
             extra_cflags=["-O2", "-DWITH_CUDA"], extra_cuda_cflags=["-O2", "-DWITH_CUDA"],
             build_directory=d, with_cuda=True, verbose=verbose)
    if "prefix_sum" in which:
        c = os.path.join(REF, "external/FRNN/external/prefix_sum")
        d = os.path.join(OUT, "ref_prefix_sum"); os.makedirs(d, exist_ok=True)
        load(name="ref_prefix_sum", sources=[os.path.join(c, "prefix_sum.cu")],
             extra_include_paths=[c], extra_cuda_cflags=["-O2"],
             build_directory=d, with_cuda=True, verbose=verbose)


if __name__ == "__main__":
    build(tuple(sys.argv[1:]) or ("frnn", "dss", "prefix_sum"), verbose=True)
