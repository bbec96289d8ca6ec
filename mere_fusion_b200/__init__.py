"""mere_fusion_b200 -- B200-native (sm_100a) audio -> face-frame hot path behind mere-fusion's
BaseReal / BaseASR plugin surface.  (The repository is "mere-fusion"; a Python package name
cannot carry the hyphen, hence the underscore.)

The compute lives in libmf_b200.so (hand-written CUDA, C ABI in include/mf_b200.h); this
package is the ctypes shim plus the host-side mirror of the reference's plugin classes.
There is no CPU or PyTorch fallback: importing the kernels without the built library raises.
"""
__version__ = "0.1.0"
