"""Oracle: FingerPrinter encoder.  TEST INFRASTRUCTURE ONLY.

Restates ``model/fp/nnfp.py`` of the reference:

* ``ConvLayer`` ``:20-83``: Conv2D(1x3, SAME) -> ELU -> LayerNorm(axis=(1,2,3)) ->
  Conv2D(3x1, SAME) -> ELU -> LayerNorm(axis=(1,2,3)); NHWC with H=F (mel), W=T (time),
* ``FingerPrinter`` ``:159-231``: 8 ConvLayers with channels ``:193`` and strides ``:194-197``,
  Flatten, ``DivEncLayer`` ``:86-156`` (128 x [Dense(8->32, elu), Dense(32->1)], BN list unused),
  ``tf.math.l2_normalize(axis=1)`` ``:229`` (x * rsqrt(max(sum x^2, 1e-12))),
* Keras ``LayerNormalization`` defaults: epsilon 1e-3, biased variance, gamma/beta of the
  normalised shape (F, T, C).

Weights arrive as a dict of numpy arrays in the exchange format of
``neural-audio-fp_b200/model/weights.py`` (conv kernels HWIO, LN params (F,T,C), stacked
div-enc tensors).  PARITY UNPINNED against TensorFlow itself; convolution / layer-norm are
cross-checked against torch.nn.functional in tests/test_oracle_encoder.py.
"""
from __future__ import annotations

import numpy as np

FRONT_HIDDEN_CH = [128, 128, 256, 256, 512, 512, 1024, 1024]          # nnfp.py:193
FRONT_STRIDES = [[(1, 2), (2, 1)], [(1, 2), (2, 1)], [(1, 2), (2, 1)], [(1, 2), (2, 1)],
                 [(1, 1), (2, 1)], [(1, 2), (2, 1)], [(1, 1), (2, 1)], [(1, 2), (2, 1)]]  # :194-197
LN_EPS = 1e-3        # Keras LayerNormalization default
L2_EPS = 1e-12       # tf.math.l2_normalize default


def same_pad(n_in, k, s):
    """TF 'SAME': out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0); lo = total//2."""
    n_out = -(-n_in // s)
    total = max((n_out - 1) * s + k - n_in, 0)
    lo = total // 2
    return n_out, lo, total - lo


def layer_table(input_shape=(256, 32, 1)):
    """[(name, kernel(kh,kw), stride(sf,st), in(F,T,C), out(F,T,C))] for the 16 convolutions."""
    F, T, C = input_shape
    rows = []
    for i, (ch, (s_a, s_b)) in enumerate(zip(FRONT_HIDDEN_CH, FRONT_STRIDES)):
        Fo, _, _ = same_pad(F, 1, s_a[0])
        To, _, _ = same_pad(T, 3, s_a[1])
        rows.append((f"conv{i}_a", (1, 3), s_a, (F, T, C), (Fo, To, ch)))
        F, T, C = Fo, To, ch
        Fo, _, _ = same_pad(F, 3, s_b[0])
        To, _, _ = same_pad(T, 1, s_b[1])
        rows.append((f"conv{i}_b", (3, 1), s_b, (F, T, C), (Fo, To, ch)))
        F, T, C = Fo, To, ch
    return rows


def param_count(input_shape=(256, 32, 1), emb=128, unit=(32, 1)):
    n = 0
    for _, (kh, kw), _, (_, _, ci), (fo, to, co) in layer_table(input_shape):
        n += kh * kw * ci * co + co + 2 * fo * to * co
    last = layer_table(input_shape)[-1][4]
    flat = last[0] * last[1] * last[2]
    s = flat // emb
    n += emb * (s * unit[0] + unit[0] + unit[0] * unit[1] + unit[1])
    return n


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def conv2d_same_nhwc(x, w, b, stride):
    """x (B,F,T,Ci), w (kh,kw,Ci,Co) HWIO, TF SAME padding; plain numpy (einsum per tap)."""
    B, F, T, Ci = x.shape
    kh, kw, _, Co = w.shape
    sf, st = stride
    Fo, flo, fhi = same_pad(F, kh, sf)
    To, tlo, thi = same_pad(T, kw, st)
    xp = np.pad(x, ((0, 0), (flo, fhi), (tlo, thi), (0, 0)))
    out = np.zeros((B, Fo, To, Co), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (Fo - 1) * sf + 1:sf, j:j + (To - 1) * st + 1:st, :]
            out += np.tensordot(patch, w[i, j], axes=([3], [0]))
    return out + b


def layer_norm_ftc(x, gamma, beta, eps=LN_EPS):
    """Keras LayerNormalization(axis=(1,2,3)): per-sample mean / biased variance over (F,T,C)."""
    mu = x.mean(axis=(1, 2, 3), keepdims=True)
    var = ((x - mu) ** 2).mean(axis=(1, 2, 3), keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * gamma[None] + beta[None]


def front_conv(x, weights, return_all=False):
    """x (B,256,32,1) -> (B,1024).  ``return_all`` also returns every post-LN activation."""
    acts = []
    for i, (s_a, s_b) in enumerate(FRONT_STRIDES):
        x = conv2d_same_nhwc(x, weights[f"conv{i}_a_w"], weights[f"conv{i}_a_b"], s_a)
        x = layer_norm_ftc(_elu(x), weights[f"ln{i}_a_g"], weights[f"ln{i}_a_b"])
        acts.append(x)
        x = conv2d_same_nhwc(x, weights[f"conv{i}_b_w"], weights[f"conv{i}_b_b"], s_b)
        x = layer_norm_ftc(_elu(x), weights[f"ln{i}_b_g"], weights[f"ln{i}_b_b"])
        acts.append(x)
    flat = x.reshape(x.shape[0], -1)
    return (flat, acts) if return_all else flat


def div_enc(x, weights):
    """(B, 1024) -> (B, 128): slice i = features 8i..8i+7 (``nnfp.py:155``), Dense(32, elu), Dense(1)."""
    w1, b1, w2, b2 = (weights[k] for k in ("div_w1", "div_b1", "div_w2", "div_b2"))
    q, s, _ = w1.shape
    xs = x.reshape(x.shape[0], q, s)
    h = _elu(np.einsum("bqs,qsu->bqu", xs, w1) + b1[None])
    return np.einsum("bqu,quo->bqo", h, w2)[..., 0] + b2[None, :, 0]


def l2_normalize(x, eps=L2_EPS):
    ss = (x * x).sum(axis=1, keepdims=True)
    return x / np.sqrt(np.maximum(ss, x.dtype.type(eps)))


def fingerprinter(mel, weights, dtype=np.float64, return_all=False):
    """``FingerPrinter.call`` (``nnfp.py:224-231``): mel (B,256,32,1) -> (B,128) unit-norm."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in weights.items()}
    x = np.asarray(mel, dtype=dtype)
    if return_all:
        flat, acts = front_conv(x, w, True)
        return l2_normalize(div_enc(flat, w)), acts
    return l2_normalize(div_enc(front_conv(x, w), w))
