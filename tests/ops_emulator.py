"""TEST INFRASTRUCTURE: a torch (CPU, any dtype) statement of the operator contracts of ``include/reptext_rt.h`` that the
VAE host code (``reptext_b200/vae.py``) drives - ``rt_gemm`` with its BIAS / GATE_RESID / SCALE_MASK epilogues and its
conv mode, GroupNorm, upsampling, row softmax, im2col, layout changes, posterior sampling.  ``tests/test_vae_host_logic.py``
monkeypatches ``reptext_b200.ops`` with these to check the host logic (weight packing, padding, bias folding, block order)
against the oracle WITHOUT a GPU.  Never imported by the product."""
import torch
import torch.nn.functional as F

from reptext_b200 import _lib as L


def _im2col(x, H, W, C, Ho, Wo, stride, pad_lo, Kp):
    B = x.shape[0]
    img = x.view(B, H, W, -1)[..., :C]
    pad_hi_y = max(0, (Ho - 1) * stride + 2 - pad_lo - (H - 1))
    pad_hi_x = max(0, (Wo - 1) * stride + 2 - pad_lo - (W - 1))
    p = F.pad(img, (0, 0, pad_lo, pad_hi_x, pad_lo, pad_hi_y))
    cols = []
    for ky in range(3):
        for kx in range(3):
            cols.append(p[:, ky:ky + (Ho - 1) * stride + 1:stride, kx:kx + (Wo - 1) * stride + 1:stride, :])
    out = torch.zeros(B, Ho * Wo, Kp, dtype=x.dtype)
    out[..., :9 * C] = torch.cat(cols, dim=-1).reshape(B, Ho * Wo, 9 * C)
    return out


def gemm(problems, batch, dtype, rope=None, head_dim=0, impl=0, sp_out=None, sp_cols=0, sp_row0=0):
    for p in problems:
        A = p.A
        m = p.m_rows if p.m_rows is not None else A.shape[1] - p.a_row0
        if p.conv_hw is not None:
            H, W = p.conv_hw
            C = p.conv_c if p.conv_c is not None else A.shape[2]
            cp = (C + 63) // 64 * 64
            a = torch.zeros(A.shape[0], H * W, 9, cp, dtype=A.dtype)
            a[..., :C] = _im2col(A, H, W, C, H, W, 1, 1, 9 * C).view(A.shape[0], H * W, 9, C)
            a = a.view(A.shape[0], H * W, 9 * cp)
            assert p.K == 9 * cp
        else:
            K = p.K if p.K is not None else A.shape[2]
            a = A[:, p.a_row0:p.a_row0 + m, :K]
        a = a.double()
        if a.shape[0] == 1 and batch > 1:
            a = a.expand(batch, -1, -1)
        for s in p.segs:
            n = s.W.shape[0]
            assert s.W.is_contiguous() and n % 64 == 0, "segment rows must be a multiple of 64"
            acc = a @ s.W.double().t()
            if s.bias is not None:
                acc = acc + s.bias.double()
            view = s.out[:, p.out_row0:p.out_row0 + m, s.out_col0:s.out_col0 + n]
            if s.mode == L.EPI_BIAS:
                r = acc
            elif s.mode == L.EPI_GELU:
                r = F.gelu(acc, approximate="tanh")
            elif s.mode == L.EPI_GATE_RESID:
                g = p.gate.double()[:, None, :] if p.gate is not None else 1.0
                r = view.double() + g * acc
            elif s.mode == L.EPI_SCALE_MASK:
                r = acc * p.scale
                if p.mask is not None:
                    r = r * p.mask.double()[None, :, None]
                if p.accumulate:
                    r = r + view.double()
            else:
                raise NotImplementedError(s.mode)
            view.copy_(r.to(s.out.dtype))


def groupnorm_nhwc(x, groups, gamma, beta, eps=1e-6, silu=False, out=None):
    y = F.group_norm(x.double().transpose(1, 2), groups, gamma.double(), beta.double(), eps=eps)
    if silu:
        y = F.silu(y)
    return y.transpose(1, 2).to(x.dtype).contiguous()


def upsample_nearest2x_nhwc(x, hw):
    B, HW, C = x.shape
    v = x.view(B, hw[0], hw[1], C)
    return v.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).reshape(B, 4 * HW, C).contiguous()


def softmax_rows_(x):
    x.copy_(torch.softmax(x.double(), dim=-1).to(x.dtype))
    return x


def softmax_rows_f32(x, out):
    out.copy_(torch.softmax(x.double(), dim=-1).to(out.dtype))
    return out


def im2col3x3_nhwc(x, hw, C_used, out_hw, stride, pad_lo, Kp=None):
    Kp = Kp or (9 * C_used + 7) // 8 * 8
    return _im2col(x, hw[0], hw[1], C_used, out_hw[0], out_hw[1], stride, pad_lo, Kp)


def nchw_to_nhwc(x, c_pad, dtype=None):
    B, C, H, W = x.shape
    out = torch.zeros(B, H * W, c_pad, dtype=dtype or x.dtype)
    out[..., :C] = x.permute(0, 2, 3, 1).reshape(B, H * W, C)
    return out


def nhwc_to_nchw(x, hw, C_used, dtype):
    B = x.shape[0]
    return x[..., :C_used].reshape(B, hw[0], hw[1], C_used).permute(0, 3, 1, 2).to(dtype).contiguous()


def vae_posterior_sample(moments, hw, latent_channels, noise, dtype):
    m = nhwc_to_nchw(moments, hw, 2 * latent_channels, torch.float64)
    z = m[:, :latent_channels]
    if noise is not None:
        z = z + torch.exp(0.5 * m[:, latent_channels:].clamp(-30, 20)) * noise.double()
    return z.to(dtype)


def norm_rows(x, weight, bias, eps, subtract_mean):
    t = x.double()
    if subtract_mean:
        t = t - t.mean(-1, keepdim=True)
    y = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + eps) * weight.double()
    if bias is not None:
        y = y + bias.double()
    return y.to(x.dtype)


def text_attention(qkv, heads, scale, rel_bias=None, causal=False):
    B, S, W3 = qkv.shape
    D = W3 // 3
    q, k, v = [t.double().view(B, S, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=-1)]
    s = scale * (q @ k.transpose(-1, -2))
    i = torch.arange(S)
    if rel_bias is not None:
        s = s + rel_bias.double()[:, (i[None, :] - i[:, None]) + S - 1][None]
    if causal:
        s = s + torch.full((S, S), float("-inf"), dtype=torch.float64).triu(1)
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, D).to(qkv.dtype)


def glu_act(x, F_out, kind):
    a = x[..., :F_out].double()
    r = a * x[..., F_out:2 * F_out].double() if kind == 0 else a * torch.sigmoid(1.702 * a)
    return r.to(x.dtype)


def embedding(table, ids, pos_table=None):
    if int(ids.min()) < 0 or int(ids.max()) >= table.shape[0]:
        raise IndexError("token id outside the embedding table")
    out = table[ids]
    if pos_table is not None:
        out = out + pos_table[: ids.shape[1]]
    return out


def install(monkeypatch, ops_module):
    for name in ("gemm", "groupnorm_nhwc", "upsample_nearest2x_nhwc", "softmax_rows_", "softmax_rows_f32", "im2col3x3_nhwc", "nchw_to_nhwc",
                 "nhwc_to_nchw", "vae_posterior_sample", "norm_rows", "text_attention", "glu_act", "embedding"):
        monkeypatch.setattr(ops_module, name, globals()[name])
