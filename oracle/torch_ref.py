"""Oracle, timed form: the extractor on torch-CPU fp32 with all host threads -- what the reference's
TensorFlow-CPU ``generate`` does (oneDNN convolutions, fused layer norm), for the ``cpu_baseline`` of fingerprint
generation.  TEST INFRASTRUCTURE ONLY.  Same arithmetic as oracle/melspec.py + oracle/fingerprinter.py (which stay the
parity oracles, fp64); tests/test_oracle_encoder.py checks the two against each other."""
from __future__ import annotations

import numpy as np

from .fingerprinter import FRONT_STRIDES, L2_EPS, LN_EPS, same_pad


def _prep(weights):
    import torch
    w = {}
    for i in range(len(FRONT_STRIDES)):
        for ab in "ab":
            k = torch.from_numpy(np.asarray(weights[f"conv{i}_{ab}_w"], np.float32))            # HWIO
            w[f"conv{i}_{ab}_w"] = k.permute(3, 2, 0, 1).contiguous()                           # OIHW
            w[f"conv{i}_{ab}_b"] = torch.from_numpy(np.asarray(weights[f"conv{i}_{ab}_b"], np.float32))
            for gb in "gb":
                p = torch.from_numpy(np.asarray(weights[f"ln{i}_{ab}_{gb}"], np.float32))       # (F, T, C)
                w[f"ln{i}_{ab}_{gb}"] = p.permute(2, 0, 1).contiguous()                         # (C, F, T)
    for k in ("div_w1", "div_b1", "div_w2", "div_b2"):
        w[k] = torch.from_numpy(np.asarray(weights[k], np.float32))
    return w


class TorchFingerPrinter:
    """nnfp.py:159-231 on torch-CPU: call with mel (B, 256, 32, 1) float32 -> (B, 128) float32."""

    def __init__(self, weights):
        self.w = _prep(weights)

    def __call__(self, mel):
        import torch
        import torch.nn.functional as F
        w = self.w
        with torch.no_grad():
            x = torch.from_numpy(np.ascontiguousarray(mel, np.float32)).permute(0, 3, 1, 2)     # (B, C=1, F, T)
            for i, (s_a, s_b) in enumerate(FRONT_STRIDES):
                for ab, (kh, kw), st in (("a", (1, 3), s_a), ("b", (3, 1), s_b)):
                    _, flo, fhi = same_pad(x.shape[2], kh, st[0])
                    _, tlo, thi = same_pad(x.shape[3], kw, st[1])
                    x = F.pad(x, (tlo, thi, flo, fhi))                                          # TF 'SAME' (asymmetric)
                    x = F.elu(F.conv2d(x, w[f"conv{i}_{ab}_w"], w[f"conv{i}_{ab}_b"], stride=st))
                    x = F.layer_norm(x, x.shape[1:], w[f"ln{i}_{ab}_g"], w[f"ln{i}_{ab}_b"], eps=LN_EPS)
            flat = x.reshape(x.shape[0], -1)                                                    # (B, 1024): F = T = 1
            xs = flat.reshape(flat.shape[0], 128, 8)
            h = F.elu(torch.einsum("bqs,qsu->bqu", xs, w["div_w1"]) + w["div_b1"][None])
            y = torch.einsum("bqu,quo->bqo", h, w["div_w2"])[..., 0] + w["div_b2"][None, :, 0]
            y = y / torch.sqrt(torch.clamp((y * y).sum(1, keepdim=True), min=L2_EPS))
            return y.numpy()


def melspec_torch(x, group_size=None):
    """melspectrogram.py:59-112 on torch-CPU fp32: x (B, 1, 8000) -> (B, 256, 32, 1); consecutive groups of
    ``group_size`` rows share the batch-global max."""
    import torch
    from .melspec import mel_filterbank
    with torch.no_grad():
        xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).reshape(len(x), -1)
        xt = torch.nn.functional.pad(xt, (512, 512))
        spec = torch.stft(xt, 1024, 256, 1024, window=torch.hann_window(1024, periodic=True), center=False,
                          return_complex=True).abs()                                             # (B, 513, 32)
        fb = torch.from_numpy(np.ascontiguousarray(mel_filterbank().T, np.float32))                 # (513, 256)
        mel = torch.einsum("bft,fm->bmt", spec, fb)
        y = torch.log10(torch.clamp(mel + 0.06, min=1e-10))
        g = int(group_size) if group_size else max(len(x), 1)
        for s in range(0, len(x), g):
            y[s:s + g] = torch.clamp(y[s:s + g] - y[s:s + g].max(), min=-80.0)
        return y[..., None].numpy()
