"""Loader for the REFERENCE's own compiled ErNeRF CUDA extensions (oracle/_ref/, built by
oracle/build_ref.py from the sources under /root/reference).  Test infrastructure only."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("_raymarching_face", "_grid_encoder", "_sh_encoder", "_freqencoder")


def available():
    return all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", n, n + ".so")) for n in NAMES)


def load():
    import torch  # noqa: F401  (libtorch must be resident before the pybind modules load)
    mods = {}
    for n in NAMES:
        spec = importlib.util.spec_from_file_location(n, os.path.join(ROOT, "oracle", "_ref", n, n + ".so"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[n] = m
    return mods
