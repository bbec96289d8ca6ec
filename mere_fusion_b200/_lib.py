"""ctypes binding of libmf_b200.so (include/mf_b200.h).  Fails loudly when the library is absent."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MF_B200_LIB") or os.path.join(HERE, "libmf_b200.so")   # MF_B200_LIB: an experiment build of the same library

MF_ERNERF_HEAD_LEVELS = 12
MF_ERNERF_TORSO_LEVELS = 16

c_f = ctypes.c_float
c_u32 = ctypes.c_uint32
c_i32 = ctypes.c_int32
c_vp = ctypes.c_void_p


class MfErnerfCfg(ctypes.Structure):
    _fields_ = [("bound", c_f), ("min_near", c_f), ("dt_gamma", c_f), ("T_thresh", c_f),
                ("density_thresh_torso", c_f), ("torso_shrink", c_f), ("max_steps", c_u32),
                ("cascade", c_u32), ("grid_size", c_u32), ("smooth_lips", c_u32),
                ("head_log2_scale", c_f), ("head_base", c_u32),
                ("head_offsets", c_i32 * (MF_ERNERF_HEAD_LEVELS + 1)),
                ("torso_log2_scale", c_f), ("torso_base", c_u32),
                ("torso_offsets", c_i32 * (MF_ERNERF_TORSO_LEVELS + 1)), ("audio_in_dim", c_u32)]


class MfErnerfFrame(ctypes.Structure):
    _fields_ = [("pose", ctypes.POINTER(c_f)), ("fx", c_f), ("fy", c_f), ("cx", c_f), ("cy", c_f),
                ("H", c_i32), ("W", c_i32), ("auds", c_vp), ("enc_a", c_vp), ("eye", c_f),
                ("bg_color", c_vp), ("rays_o", c_vp), ("rays_d", c_vp), ("bg_coords", c_vp),
                ("n_rays", c_i32), ("outH", c_i32), ("outW", c_i32), ("out_image_f32", c_vp)]


class MfErnerfDebug(ctypes.Structure):
    _fields_ = [("nears", c_vp), ("fars", c_vp), ("round_info", c_vp), ("weights_sum", c_vp),
                ("image_head", c_vp), ("enc_a", c_vp), ("torso_mask", c_vp)]


_lib = None


class MfError(RuntimeError):
    pass


def lib():
    """The loaded library.  There is deliberately no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MfError(f"{LIB_PATH} is missing: build it with `python -m mere_fusion_b200.build` "
                          "(there is no CPU / PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.mf_version.restype = ctypes.c_int
        L.mf_create.argtypes = [ctypes.c_int, ctypes.POINTER(c_vp)]
        L.mf_destroy.argtypes = [c_vp]
        L.mf_destroy.restype = None
        L.mf_last_error.argtypes = [c_vp]
        L.mf_last_error.restype = ctypes.c_char_p
        L.mf_ernerf_load.argtypes = [c_vp, c_vp, ctypes.c_size_t, ctypes.POINTER(MfErnerfCfg)]
        L.mf_ernerf_render.argtypes = [c_vp, ctypes.POINTER(MfErnerfFrame), c_vp,
                                       ctypes.POINTER(MfErnerfDebug), c_vp]
        L.mf_ernerf_render_batch.argtypes = [ctypes.POINTER(c_vp), ctypes.POINTER(MfErnerfFrame), ctypes.POINTER(c_vp), ctypes.c_int, c_vp]
        L.mf_ernerf_reset_state.argtypes = [c_vp]
        L.mf_ernerf_encode_audio.argtypes = [c_vp, c_vp, c_vp, c_vp]
        L.mf_ernerf_last_launches.argtypes = [c_vp]
        L.mf_ernerf_profile.argtypes = [c_vp, ctypes.c_int]
        L.mf_ernerf_last_head_ms.argtypes = [c_vp, ctypes.POINTER(c_f), ctypes.POINTER(ctypes.c_int64)]
        L.mf_ernerf_blob_layout.argtypes = [ctypes.POINTER(c_i32), ctypes.c_int]
        L.mf_near_far_from_aabb.argtypes = [c_vp, c_vp, c_vp, c_vp, c_u32, c_f, c_vp, c_vp, c_vp]
        L.mf_march_rays.argtypes = [c_vp, c_u32, c_u32, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_u32, c_u32, c_u32,
                                    c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
        L.mf_composite_rays_triplane.argtypes = [c_vp, c_u32, c_u32, c_f] + [c_vp] * 15
        L.mf_grid_encode_forward.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_u32, c_u32, c_u32, c_u32, c_f,
                                             c_u32, c_u32, ctypes.c_int, ctypes.c_int, c_vp]
        L.mf_grid_level_scales.argtypes = [c_vp, c_f, c_u32, c_u32, ctypes.POINTER(c_f)]
        L.mf_sh_encode_forward.argtypes = [c_vp, c_vp, c_vp, c_u32, c_u32, c_u32, c_vp]
        L.mf_freq_encode_forward.argtypes = [c_vp, c_vp, c_u32, c_u32, c_u32, c_u32, c_vp, c_vp]
        L.mf_wav2lip_load.argtypes = [c_vp, c_vp, ctypes.c_size_t, ctypes.c_int]
        L.mf_wav2lip_forward.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, c_vp]
        L.mf_wav2lip_last_launches.argtypes = [c_vp]
        L.mf_wav2lip_profile.argtypes = [c_vp, ctypes.c_int]
        L.mf_wav2lip_last_op_ms.argtypes = [c_vp, ctypes.POINTER(c_f)]
        L.mf_convnet_debug_run.argtypes = [c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp]
        L.mf_musetalk_forward.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, c_vp]
        L.mf_convnet_debug_set.argtypes = [c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp]
        L.mf_paste_resize_u8.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int,
                                         ctypes.c_int, ctypes.POINTER(c_i32), c_vp, c_vp]
        L.mf_wav2vec2_logits.argtypes = [c_vp, c_vp, ctypes.c_int, c_vp, c_vp]
        L.mf_wav2vec2_logits_batch.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp]
        L.mf_debug_w2v_phase_ns.argtypes = [c_vp, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
        L.mf_whisper_features.argtypes = [c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp]
        L.mf_wav2lip_mel_chunks.argtypes = [c_vp, c_vp, ctypes.c_int, c_vp, ctypes.POINTER(c_i32), ctypes.c_int, c_vp, c_vp]
        L.mf_paste_blend_u8.argtypes = [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int,
                                        ctypes.c_int, ctypes.POINTER(c_i32), c_vp, ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_int64), c_vp, c_vp]
        _lib = L
    return _lib


def check(ctx, rc, what):
    if rc != 0:
        msg = lib().mf_last_error(ctx)
        raise MfError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


class Context:
    """One mf_ctx per (GPU, session)."""

    def __init__(self, device=0):
        self._h = c_vp()
        rc = lib().mf_create(int(device), ctypes.byref(self._h))
        if rc != 0:
            raise MfError(f"mf_create(device={device}) failed ({rc}): an sm_100 (B200) GPU is required; "
                          "there is no CPU fallback")
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().mf_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
