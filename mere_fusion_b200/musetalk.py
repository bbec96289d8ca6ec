"""MuseTalkEngine: torch tensors in/out around mf_musetalk_forward (include/mf_b200.h).  Replaces
pe(...) -> unet.model(...).sample -> vae.decode_latents(...) of musereal.py:99-108."""
import ctypes

import torch

from ._lib import check, lib
from .musetalk_pack import UNET_CFG, VAE_CFG, pack_musetalk
from .wav2lip import ConvNet, _ptr


class MuseTalkEngine(ConvNet):
    def __init__(self, unet_sd=None, vae_sd=None, ucfg=UNET_CFG, vcfg=VAE_CFG, max_batch=16, device=0, blob=None):
        self.flops_per_frame = None
        if blob is None:
            blob, pb = pack_musetalk(unet_sd, vae_sd, ucfg, vcfg, nominal_batch=max_batch)
            self.flops_per_frame = pb.flops_per_sample
            self.unet_flops, self.vae_flops = pb.unet_flops, pb.vae_flops
            self.n_ops = len(pb.ops)
            self.op_records = list(pb.ops)
        self.out_hw = 256
        super().__init__(blob, max_batch, device)

    def forward(self, latents, whisper, out=None, out_f32=None, stream=None):
        """latents: cuda fp16 [B,8,32,32]; whisper: cuda fp16 [B,50,384] -> u8 [B,256,256,3] BGR"""
        B = int(latents.shape[0])
        assert latents.is_cuda and whisper.is_cuda and latents.dtype == torch.float16 and whisper.dtype == torch.float16
        assert latents.is_contiguous() and whisper.is_contiguous() and whisper.shape[0] == B
        if out is None:
            out = torch.empty((B, self.out_hw, self.out_hw, 3), dtype=torch.uint8, device=self.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(self.ctx.handle, lib().mf_musetalk_forward(self.ctx.handle, _ptr(latents), _ptr(whisper), _ptr(out), _ptr(out_f32), B,
                                                         ctypes.c_void_p(s.cuda_stream)), "mf_musetalk_forward")
        return out
