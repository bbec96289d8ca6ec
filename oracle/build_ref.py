"""Recipe: compile the REFERENCE's own ErNeRF CUDA extensions for sm_100 into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.

The four extensions (raymarching, gridencoder, shencoder, freqencoder) are compiled from the
sources *where they lie* under /root/reference/ernerf/*/src (never copied into this repo), with
the flags of the reference's own backend.py files (ernerf/raymarching/backend.py:6-9,
ernerf/gridencoder/backend.py:6-9, ernerf/shencoder/backend.py:6-9,
ernerf/freqencoder/backend.py:6-10 -- freqencoder alone adds -use_fast_math).  Outputs (.so
pybind modules) land only in oracle/_ref/, which is git-ignored but travels to the GPU box.

The GPU-side tests import those modules (tests/ref_ernerf.py) to pin the C oracle and the
sm_100a kernels against the reference's own kernels.  Without /root/reference this is a no-op.
"""
import os
import sys

REF = os.environ.get("MF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

EXTS = {
    # name used by the reference's backend.py -> (subdir, sources, extra nvcc flags)
    "_raymarching_face": ("raymarching", ["raymarching.cu", "bindings.cpp"], []),
    "_grid_encoder": ("gridencoder", ["gridencoder.cu", "bindings.cpp"], []),
    "_sh_encoder": ("shencoder", ["shencoder.cu", "bindings.cpp"], []),
    "_freqencoder": ("freqencoder", ["freqencoder.cu", "bindings.cpp"], ["-use_fast_math"]),
}


def build(names=None, verbose=False):
    if not os.path.isdir(os.path.join(REF, "ernerf")):
        print(f"[oracle/build_ref] {REF} absent: nothing to build (prebuilt oracle/_ref is used as is)")
        return []
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    built = []
    for name, (sub, srcs, extra) in EXTS.items():
        if names and name not in names:
            continue
        so = os.path.join(OUT, name, name + ".so")
        if os.path.exists(so):
            built.append(so)
            continue
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        nvcc_flags = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__",
                      "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
                      "-allow-unsupported-compiler"] + extra
        load(name=name, extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=nvcc_flags,
             sources=[os.path.join(REF, "ernerf", sub, "src", s) for s in srcs],
             build_directory=bdir, verbose=verbose, is_python_module=False)
        built.append(so)
    return built


if __name__ == "__main__":
    print(build(sys.argv[1:] or None, verbose=True))
