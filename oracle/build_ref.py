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

stage_python() additionally STAGES (file copy at build time, never into the git tree) the reference's
own ErNeRF Python -- encoding.py, nerf_triplane/{network,renderer,utils,provider}.py and the four
extension wrapper packages -- into the git-ignored oracle/_ref/py/ernerf/, so that on the GPU box
(where /root/reference does not exist) tests/ref_ernerf.py can build the reference NeRFNetwork +
Trainer and run the reference's OWN full render (Trainer.test_gui_with_data -> run_cuda + run_torso,
utils.py:1191-1223, renderer.py:158-352) as the frame-level ground truth and as the same-box timed
GPU baseline of bench.py (`reference_cuda`).
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


PY_FILES = [
    "encoding.py",
    "nerf_triplane/network.py", "nerf_triplane/renderer.py", "nerf_triplane/utils.py", "nerf_triplane/provider.py",
    "raymarching/__init__.py", "raymarching/raymarching.py", "raymarching/backend.py",
    "gridencoder/__init__.py", "gridencoder/grid.py", "gridencoder/backend.py",
    "shencoder/__init__.py", "shencoder/sphere_harmonics.py", "shencoder/backend.py",
    "freqencoder/__init__.py", "freqencoder/freq.py", "freqencoder/backend.py",
]


def stage_python():
    """copy the reference's ErNeRF Python (unmodified) into oracle/_ref/py/ernerf/ -- git-ignored, travels with gpurun"""
    import shutil
    src = os.path.join(REF, "ernerf")
    if not os.path.isdir(src):
        return []
    out = []
    for rel in PY_FILES:
        dst = os.path.join(OUT, "py", "ernerf", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        s = os.path.join(src, rel)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(s):
            shutil.copyfile(s, dst)
        out.append(dst)
    # the reference's dense-head modules for the same-box cuDNN / cuBLAS baselines of bench.py (`heads.*.torch_gpu`):
    # wav2lip/models (wav2lip.py:87-125, conv.py) and the vendored Whisper (musetalk/whisper, encoder model.py:143-171)
    for sub in ("wav2lip/models", "musetalk/whisper"):
        s_dir, d_dir = os.path.join(REF, sub), os.path.join(OUT, "py", sub)
        if os.path.isdir(s_dir) and not os.path.isdir(d_dir):
            shutil.copytree(s_dir, d_dir, ignore=shutil.ignore_patterns("__pycache__"))
        out.append(d_dir)
    return out


def build(names=None, verbose=False):
    if not os.path.isdir(os.path.join(REF, "ernerf")):
        print(f"[oracle/build_ref] {REF} absent: nothing to build (prebuilt oracle/_ref is used as is)")
        return []
    stage_python()
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
