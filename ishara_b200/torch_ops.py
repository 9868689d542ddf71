"""Operator-level mirrors of the reference's vendored PyTorch modules, run on the sm_100a kernels.

The north star names `squeezeformer/{attention,convolution,modules,encoder}.py` and `conformer/conformer.py`; `get_model`
never calls them, so they are matched operator by operator: each function below takes the module's `state_dict()`
(numpy, PyTorch layouts: Linear `[out,in]`, Conv1d `[out,in/groups,k]`) plus fp32 host activations `[B,T,D]`, runs the
computation through the C-ABI op entry points (`ishara_op_gemm`, `ishara_op_dwconv`, `ishara_op_attention`,
`ishara_op_relpos_attention`, `ishara_op_time_reduce`, `ishara_op_upsample_add`, `ishara_op_conv2d_subsample`,
`ishara_op_layernorm`) with bf16 activations, and returns fp32 host arrays. Inference semantics (dropout off,
BatchNorm running statistics). No CPU fallback: every arithmetic step is a kernel launch.

  P1  relative_mha              squeezeformer/attention.py:25-110      RelativeMultiHeadAttention
  P2  mhsa_module               squeezeformer/attention.py:113-139     MultiHeadedSelfAttentionModule
  P3  rel_positional_encoding   squeezeformer/modules.py:59-108        RelPositionalEncoding (constant table, host)
  P4  feed_forward              squeezeformer/modules.py:24-56         FeedForwardModule
  P5  conv_module               squeezeformer/convolution.py:199-238   ConvModule
  P6  squeezeformer_block       squeezeformer/encoder.py:169-247       SqueezeformerBlock (post-LN, residual scaling)
  P7  time_reduction            squeezeformer/convolution.py:241-269   TimeReductionLayer (+ encoder.py:80,155 proj)
  P8  recover                   squeezeformer/modules.py:137-142       recover_resolution (+ encoder.py:157-162)
      conv2d_subsampling        squeezeformer/convolution.py:39-73     DepthwiseConv2dSubsampling
  P9  conformer_*               conformer/conformer.py:6-73            FeedForwardModule / MHSA / ConvolutionModule / ConformerBlock
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Mapping, Optional, Tuple

import numpy as np

from . import _lib
from ._dlpack import DeviceTensor

SD = Mapping[str, np.ndarray]


def to_bf16_bits(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16).reshape(a.shape)


def from_bf16_bits(u: np.ndarray) -> np.ndarray:
    return (u.astype(np.uint32) << 16).view(np.float32)


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _block_n(n: int) -> int:
    return 256 if n % 256 == 0 else (128 if n % 128 == 0 else 64)


def _sub(sd: SD, prefix: str) -> Dict[str, np.ndarray]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


class VendoredOps:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.device = device

    # ---- device helpers -------------------------------------------------------------------------
    def _act(self, a: np.ndarray, width: Optional[int] = None) -> DeviceTensor:
        """fp32 host [M, K] -> bf16 device [M, Kpad] (zero padded columns)."""
        M, K = a.shape
        Kp = width or _pad_to(K, 64)
        buf = np.zeros((M, Kp), np.uint16)
        buf[:, :K] = to_bf16_bits(a)
        return DeviceTensor((M, Kp), "bfloat16", self.device).copy_from_host(buf)

    def _f32(self, a) -> DeviceTensor:
        a = np.ascontiguousarray(a, dtype=np.float32)
        return DeviceTensor(a.shape, "float32", self.device).copy_from_host(a)

    def _host(self, t: DeviceTensor, cols: Optional[int] = None) -> np.ndarray:
        a = from_bf16_bits(t.numpy())
        return a[:, :cols] if cols is not None else a

    @staticmethod
    def _p(t):
        return C.c_void_p(t.ptr) if t is not None else None

    def gemm(self, a: DeviceTensor, w: np.ndarray, bias: Optional[np.ndarray] = None, *, act: int = 0,
             resid: Optional[DeviceTensor] = None, scale: float = 1.0, glu: bool = False) -> DeviceTensor:
        """out[M, Nout] bf16 = epilogue(a[M, Kp] @ w[N, K]^T): (+bias) -> act (1 swish, 2 relu) | GLU -> (+resid).
        `scale` multiplies weight and bias (module_factor of ResidualConnectionModule, modules.py:111-123)."""
        M, Kp = a.shape
        N, K = w.shape
        assert K <= Kp
        bn = _block_n(_pad_to(N, 64))
        Np = _pad_to(N, bn)
        w = np.asarray(w, np.float32) * scale
        b = np.zeros(Np, np.float32)
        if bias is not None:
            b[:N] = np.asarray(bias, np.float32) * scale
        perm = np.arange(Np)
        if glu:  # every bn-wide tile must hold [a-slice | matching gate-slice]; torch GLU: first half * sigmoid(second half)
            assert N == Np and N % 2 == 0
            half, hb = N // 2, bn // 2
            perm = np.concatenate([np.concatenate([np.arange(t * hb, (t + 1) * hb), half + np.arange(t * hb, (t + 1) * hb)])
                                   for t in range(N // bn)])
        wp = np.zeros((Np, Kp), np.float32)
        wp[:N, :K] = w
        wp, b = wp[perm], b[perm]
        wd = DeviceTensor((Np, Kp), "bfloat16", self.device).copy_from_host(to_bf16_bits(wp))
        bd = self._f32(b)
        nout = (N // 2) if glu else N
        assert nout % 64 == 0, "operator-level mirrors keep every width a multiple of 64"
        out = DeviceTensor((M, nout), "bfloat16", self.device)
        g = _lib.GemmArgs()
        g.a, g.wt, g.out0 = a.ptr, wd.ptr, out.ptr
        g.bias = bd.ptr
        g.resid = resid.ptr if resid is not None else None
        g.M, g.N, g.K, g.lda = M, Np, Kp, Kp
        g.nout = nout
        g.rows_per_seq = 1
        g.act = 3 if glu else act
        g.block_n = bn
        g.out_f32 = 0
        g.row_mode = 0
        if resid is not None:
            assert tuple(resid.shape) == tuple(out.shape)
        _lib.check(self.lib.ishara_op_gemm(C.byref(g), None))
        return out

    def layernorm(self, x: DeviceTensor, gamma, beta, eps: float = 1e-5) -> DeviceTensor:
        M, D = x.shape
        out = DeviceTensor((M, D), "bfloat16", self.device)
        g, b = self._f32(gamma), self._f32(beta)
        _lib.check(self.lib.ishara_op_layernorm(self._p(x), self._p(out), self._p(g), self._p(b), eps, M, D, None))
        return out

    def dwconv(self, x: DeviceTensor, B: int, T: int, w_kc: np.ndarray, bias: Optional[np.ndarray], pad_left: int,
               post: int) -> DeviceTensor:
        k, Cc = w_kc.shape
        out = DeviceTensor((B * T, Cc), "bfloat16", self.device)
        wd = self._f32(w_kc)
        bd = self._f32(bias) if bias is not None else None
        _lib.check(self.lib.ishara_op_dwconv(self._p(x), self._p(out), self._p(wd), self._p(bd), None, None, B, T, Cc, k,
                                             pad_left, post, None))
        return out

    # ---- P3 ---------------------------------------------------------------------------------------
    @staticmethod
    def rel_positional_encoding(T: int, D: int) -> np.ndarray:
        """modules.py:59-108: rows 0..T-1 = PE(+(T-1-r)) (the flipped positive half), rows T..2T-2 = PE(-(r-T+1));
        interleaved sin/cos. Returns [2T-1, D] float32 — a constant table, computed once on the host."""
        position = np.arange(0, T, dtype=np.float32)[:, None]
        div = np.exp(np.arange(0, D, 2, dtype=np.float32) * np.float32(-(math.log(10000.0) / D)))
        pos, neg = np.zeros((T, D), np.float32), np.zeros((T, D), np.float32)
        pos[:, 0::2], pos[:, 1::2] = np.sin(position * div), np.cos(position * div)
        neg[:, 0::2], neg[:, 1::2] = np.sin(-position * div), np.cos(-position * div)
        return np.concatenate([pos[::-1], neg[1:]], axis=0)

    # ---- P1 / P2 ----------------------------------------------------------------------------------
    def _qkv_interleaved(self, wq, bq, wk, bk, wv, bv, H: int):
        D = wq.shape[0]
        dh = D // H
        w = np.concatenate([np.concatenate([wq[h * dh:(h + 1) * dh], wk[h * dh:(h + 1) * dh], wv[h * dh:(h + 1) * dh]])
                            for h in range(H)])
        b = np.concatenate([np.concatenate([bq[h * dh:(h + 1) * dh], bk[h * dh:(h + 1) * dh], bv[h * dh:(h + 1) * dh]])
                            for h in range(H)])
        return w, b

    def relative_mha_dev(self, x: DeviceTensor, B: int, T: int, sd: SD, pos_emb: np.ndarray, H: int,
                         mask: Optional[np.ndarray] = None, resid: Optional[DeviceTensor] = None) -> DeviceTensor:
        D = sd["query_proj.weight"].shape[0]
        dh = D // H
        w, b = self._qkv_interleaved(sd["query_proj.weight"], sd["query_proj.bias"], sd["key_proj.weight"],
                                     sd["key_proj.bias"], sd["value_proj.weight"], sd["value_proj.bias"], H)
        qkv = self.gemm(x, w, b)                                               # [B*T, 3D] per-head [q|k|v]
        p = self.gemm(self._act(pos_emb.reshape(-1, D)), sd["pos_proj.weight"])  # projected once, not per batch
        u, v = self._f32(sd["u_bias"].reshape(-1)), self._f32(sd["v_bias"].reshape(-1))
        km = None
        if mask is not None:                                                   # reference: True = masked, [B,1,T]
            keep = (~np.asarray(mask, bool).reshape(B, T)).astype(np.uint8)
            km = DeviceTensor((B, T), "uint8", self.device).copy_from_host(keep)
        ctx = DeviceTensor((B * T, D), "bfloat16", self.device)
        _lib.check(self.lib.ishara_op_relpos_attention(self._p(qkv), self._p(p), self._p(u), self._p(v), self._p(ctx),
                                                       self._p(km), B, T, H, dh, 1.0 / math.sqrt(dh), None))
        return self.gemm(ctx, sd["out_proj.weight"], sd["out_proj.bias"], resid=resid)

    def relative_mha(self, x: np.ndarray, sd: SD, pos_emb: np.ndarray, H: int, mask=None) -> np.ndarray:
        B, T, D = x.shape
        return self._host(self.relative_mha_dev(self._act(x.reshape(B * T, D)), B, T, sd, pos_emb, H, mask), D).reshape(B, T, D)

    def mhsa_module(self, x: np.ndarray, sd: SD, H: int, mask=None) -> np.ndarray:
        B, T, D = x.shape
        return self.relative_mha(x, _sub(sd, "attention."), self.rel_positional_encoding(T, D), H, mask)

    # ---- P4 ---------------------------------------------------------------------------------------
    def feed_forward_dev(self, x: DeviceTensor, sd: SD, *, resid: Optional[DeviceTensor] = None, scale: float = 1.0):
        h = self.gemm(x, sd["sequential.0.weight"], sd["sequential.0.bias"], act=1)
        return self.gemm(h, sd["sequential.3.weight"], sd["sequential.3.bias"], resid=resid, scale=scale)

    def feed_forward(self, x: np.ndarray, sd: SD) -> np.ndarray:
        B, T, D = x.shape
        return self._host(self.feed_forward_dev(self._act(x.reshape(B * T, D)), sd), D).reshape(B, T, D)

    # ---- P5 ---------------------------------------------------------------------------------------
    def conv_module_dev(self, x: DeviceTensor, B: int, T: int, sd: SD, *, resid: Optional[DeviceTensor] = None):
        """pw1 -> GLU -> depthwise k 'same' (no bias) -> BatchNorm1d (eps 1e-5, running stats, folded) -> Swish -> pw2."""
        w1, b1 = sd["sequential.1.conv.weight"][:, :, 0], sd["sequential.1.conv.bias"]
        h = self.gemm(x, w1, b1, glu=True)
        wd = sd["sequential.3.conv.weight"][:, 0, :]                          # [C, k]
        k = wd.shape[1]
        s = sd["sequential.4.weight"] / np.sqrt(sd["sequential.4.running_var"] + 1e-5)
        o = sd["sequential.4.bias"] - sd["sequential.4.running_mean"] * s
        h = self.dwconv(h, B, T, (wd * s[:, None]).T.copy(), o, (k - 1) // 2, post=1)
        return self.gemm(h, sd["sequential.6.conv.weight"][:, :, 0], sd["sequential.6.conv.bias"], resid=resid)

    def conv_module(self, x: np.ndarray, sd: SD) -> np.ndarray:
        B, T, D = x.shape
        return self._host(self.conv_module_dev(self._act(x.reshape(B * T, D)), B, T, sd), D).reshape(B, T, D)

    # ---- P6 ---------------------------------------------------------------------------------------
    def squeezeformer_block(self, x: np.ndarray, sd: SD, H: int, half_step_residual: bool) -> np.ndarray:
        """encoder.py:205-244: LN(x + MHSA(x)); LN(x + f*FFN(x)); LN(x + Conv(x)); LN(x + f*FFN(x)); LN eps 1e-5."""
        B, T, D = x.shape
        f = 0.5 if half_step_residual else 1.0
        h = self._act(x.reshape(B * T, D))
        att = self.relative_mha_dev(h, B, T, _sub(sd, "sequential.0.module.attention."), self.rel_positional_encoding(T, D), H,
                                    resid=h)                                  # residual folded into out_proj's epilogue
        h = self.layernorm(att, sd["sequential.1.weight"], sd["sequential.1.bias"])
        h = self.layernorm(self.feed_forward_dev(h, _sub(sd, "sequential.2.module."), resid=h, scale=f),
                           sd["sequential.3.weight"], sd["sequential.3.bias"])
        h = self.layernorm(self.conv_module_dev(h, B, T, _sub(sd, "sequential.4.module."), resid=h),
                           sd["sequential.5.weight"], sd["sequential.5.bias"])
        h = self.layernorm(self.feed_forward_dev(h, _sub(sd, "sequential.6.module."), resid=h, scale=f),
                           sd["sequential.7.weight"], sd["sequential.7.bias"])
        return self._host(h, D).reshape(B, T, D)

    # ---- P7 / P8 ----------------------------------------------------------------------------------
    def time_reduction(self, x: np.ndarray, sd: SD, lengths: np.ndarray, proj_sd: Optional[SD] = None):
        B, T, D = x.shape
        T2, D2 = (T - 3) // 2 + 1, (D - 3) // 2 + 1
        ldo = _pad_to(D2, 64)
        xd = self._act(x.reshape(B * T, D), width=D)
        out = DeviceTensor((B * T2, ldo), "bfloat16", self.device)
        w9 = np.ascontiguousarray(sd["sequential.0.conv.weight"].reshape(9), np.float32)
        _lib.check(self.lib.ishara_op_time_reduce(self._p(xd), self._p(out), w9.ctypes.data_as(C.c_void_p),
                                                  float(sd["sequential.0.conv.bias"][0]), B, T, D, ldo, None))
        red = self._host(out, D2).reshape(B, T2, D2)
        new_len = (np.asarray(lengths) >> 1) - 1                                  # convolution.py:266-267
        if proj_sd is None:
            return red, new_len
        y = self.gemm(out, proj_sd["weight"], proj_sd["bias"])
        return red, new_len, self._host(y, proj_sd["weight"].shape[0]).reshape(B, T2, -1)

    def recover(self, small: np.ndarray, recover_tensor: np.ndarray, sd: SD) -> np.ndarray:
        """encoder.py:157-162: recover_resolution (x2 repeat) -> Linear -> += recover_tensor[:, :len]."""
        B, T2, D = small.shape
        T = recover_tensor.shape[1]
        y = self.gemm(self._act(small.reshape(B * T2, D)), sd["weight"], sd["bias"])   # Linear commutes with the repeat
        rec = self._act(recover_tensor.reshape(B * T, D), width=D)
        out = DeviceTensor((B * 2 * T2, D), "bfloat16", self.device)
        _lib.check(self.lib.ishara_op_upsample_add(self._p(y), self._p(rec), self._p(out), B, T2, T, D, None))
        return self._host(out, D).reshape(B, 2 * T2, D)

    def conv2d_subsampling(self, x: np.ndarray, sd: SD, lengths: np.ndarray):
        B, T, F = x.shape
        Cc = sd["sequential.0.weight"].shape[0]
        T4, F4 = ((T - 3) // 2 + 1 - 3) // 2 + 1, ((F - 3) // 2 + 1 - 3) // 2 + 1
        ldo = _pad_to(Cc * F4, 64)
        out = DeviceTensor((B * T4, ldo), "bfloat16", self.device)
        xd = self._f32(x)
        w1, b1 = self._f32(sd["sequential.0.weight"].reshape(Cc, 9)), self._f32(sd["sequential.0.bias"])
        w2, b2 = self._f32(sd["sequential.2.conv.weight"].reshape(Cc, 9)), self._f32(sd["sequential.2.conv.bias"])
        _lib.check(self.lib.ishara_op_conv2d_subsample(self._p(xd), self._p(out), self._p(w1), self._p(b1), self._p(w2),
                                                       self._p(b2), B, T, F, Cc, ldo, None))
        return self._host(out, Cc * F4).reshape(B, T4, Cc * F4), (np.asarray(lengths) >> 2) - 1

    # ---- P9: conformer/conformer.py ---------------------------------------------------------------
    def conformer_ffn_dev(self, x: DeviceTensor, sd: SD) -> DeviceTensor:       # :6-22, LN(FFN(x) + x)
        h = self.gemm(x, sd["linear1.weight"], sd["linear1.bias"], act=1)
        h = self.gemm(h, sd["linear2.weight"], sd["linear2.bias"], resid=x)
        return self.layernorm(h, sd["layer_norm.weight"], sd["layer_norm.bias"])

    def conformer_mhsa_dev(self, x: DeviceTensor, B: int, T: int, sd: SD, H: int) -> DeviceTensor:   # :24-35
        D = x.shape[1]
        dh = D // H
        wi, bi = sd["attention.in_proj_weight"], sd["attention.in_proj_bias"]   # [3D, D] = [q | k | v] blocks
        w, b = self._qkv_interleaved(wi[:D], bi[:D], wi[D:2 * D], bi[D:2 * D], wi[2 * D:], bi[2 * D:], H)
        qkv = self.gemm(x, w, b)
        ctx = DeviceTensor((B * T, D), "bfloat16", self.device)
        _lib.check(self.lib.ishara_op_attention(self._p(qkv), self._p(ctx), None, B, T, H, dh, 1.0 / math.sqrt(dh), None))
        h = self.gemm(ctx, sd["attention.out_proj.weight"], sd["attention.out_proj.bias"], resid=x)
        return self.layernorm(h, sd["layer_norm.weight"], sd["layer_norm.bias"])

    def conformer_conv_dev(self, x: DeviceTensor, B: int, T: int, sd: SD) -> DeviceTensor:           # :37-57
        h = self.gemm(x, sd["pointwise_conv1.weight"][:, :, 0], sd["pointwise_conv1.bias"], glu=True)
        wd = sd["depthwise_conv.weight"][:, 0, :]
        k = wd.shape[1]
        s = sd["batch_norm.weight"] / np.sqrt(sd["batch_norm.running_var"] + 1e-5)
        o = sd["batch_norm.bias"] - sd["batch_norm.running_mean"] * s + sd["depthwise_conv.bias"] * s
        h = self.dwconv(h, B, T, (wd * s[:, None]).T.copy(), o, k // 2, post=0)
        h = self.gemm(h, sd["pointwise_conv2.weight"][:, :, 0], sd["pointwise_conv2.bias"], resid=x)
        return self.layernorm(h, sd["layer_norm.weight"], sd["layer_norm.bias"])

    def conformer_block(self, x: np.ndarray, sd: SD, H: int, parts: bool = False):
        B, T, D = x.shape
        xd = self._act(x.reshape(B * T, D))
        res = {}
        if parts:
            res["ffn1"] = self._host(self.conformer_ffn_dev(xd, _sub(sd, "ffn1.")), D).reshape(B, T, D)
            res["attn"] = self._host(self.conformer_mhsa_dev(xd, B, T, _sub(sd, "attention."), H), D).reshape(B, T, D)
            res["conv"] = self._host(self.conformer_conv_dev(xd, B, T, _sub(sd, "conv.")), D).reshape(B, T, D)
        h = self.conformer_ffn_dev(xd, _sub(sd, "ffn1."))
        h = self.conformer_mhsa_dev(h, B, T, _sub(sd, "attention."), H)
        h = self.conformer_conv_dev(h, B, T, _sub(sd, "conv."))
        h = self.conformer_ffn_dev(h, _sub(sd, "ffn2."))
        h = self.layernorm(h, sd["layer_norm.weight"], sd["layer_norm.bias"])
        res["out"] = self._host(h, D).reshape(B, T, D)
        return res
