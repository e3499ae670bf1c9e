"""TEST DOUBLE (test infrastructure, never imported by the product package).

A torch re-statement of every method of `hdiff_b200.ops.CudaOps`, operand for operand, so that the
host logic of the path (kernel schedule of the forward and backward passes, packed weight layouts,
space-to-depth views, flat gradient buffers, data-parallel bucketing) can be tested against the
oracle on machines without a GPU.  The `-m gpu` tests do not use it; they run the CUDA library.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def to_logical(x, P):
    """physical NHWC [N,PH,PW,C] -> logical [N,H,W,P*P*C] with channel order (py, px, c)."""
    if P == 1:
        return x
    N, PH, PW, C = x.shape
    return x.reshape(N, PH // 2, 2, PW // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, PH // 2, PW // 2, 4 * C)


def from_logical(y, P):
    if P == 1:
        return y
    N, H, W, C4 = y.shape
    C = C4 // 4
    return y.reshape(N, H, W, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * H, 2 * W, C)


def _swish(x):
    return x * torch.sigmoid(x)


class EmuOps:
    name = "emu"

    def __init__(self):
        self.launches = 0
        self.tc_launches = 0
        self.use_tc = False

    # ---- convolution family ----
    def _logical_in(self, x0, x1, P_in, in_nchw):
        if in_nchw:
            if P_in == 2:       # 2x2 space-to-depth view of an NCHW image, logical channel order (py, px, c)
                N, C, PH, PW = x0.shape
                return x0.float().view(N, C, PH // 2, 2, PW // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 4 * C, PH // 2, PW // 2)
            return x0.float()
        x = x0 if x1 is None else torch.cat([x0, x1], dim=-1)
        return to_logical(x, P_in).permute(0, 3, 1, 2).float()

    def conv(self, x0, x1, P_in, w, bias, emb, res, out, P_out, N, H, W, k, in_nchw=False, out_nchw=False, alg_frac=1.0, Cout_pad=None, chan_sums=None, quiet=False):
        xin = self._logical_in(x0, x1, P_in, in_nchw)
        CinL = xin.shape[1]
        CoutL = w.numel() // (k * k * CinL)
        wt = w.float().view(CoutL, k, k, CinL).permute(0, 3, 1, 2)
        y = F.conv2d(xin, wt, None if bias is None else bias.float()[:CoutL], padding=k // 2)
        if emb is not None:
            y = y + emb.float()[:, :, None, None]
        if out_nchw:
            out.copy_(y[:, :out.shape[1]])
        else:
            y = from_logical(y.permute(0, 2, 3, 1), P_out)
            if res is not None:
                y = y + res.float()
            out.copy_(y.to(out.dtype))
        self.launches += 1
        if chan_sums is not None and not out_nchw and P_out == 1 and out.dtype == torch.bfloat16:
            o = out.double().reshape(N, -1, out.shape[-1])          # statistics of the values as stored
            chan_sums[:, :, 0] += o.sum(1)
            chan_sums[:, :, 1] += (o * o).sum(1)
            return True
        return None

    def wgrad(self, x0, x1, P_in, dy, P_dy, dw, N, H, W, k, dtype, in_nchw=False, dy_nchw=False, workspace=None, alg_frac=1.0, quiet=False):
        xin = self._logical_in(x0, x1, P_in, in_nchw)
        g = dy.float() if dy_nchw else to_logical(dy, P_dy).permute(0, 3, 1, 2).float()
        CinL, CoutL = xin.shape[1], g.shape[1]
        gw = torch.nn.grad.conv2d_weight(xin, (CoutL, CinL, k, k), g, padding=k // 2)
        dw.copy_(gw.permute(0, 2, 3, 1).reshape(-1))
        self.launches += 1

    def pad_nchw(self, x, out):
        out.zero_()
        out[..., :x.shape[1]] = x.permute(0, 2, 3, 1).to(out.dtype)
        self.launches += 1

    # ---- attention ----
    def attn_fwd(self, qkv, out, lse, N, S, C):
        q, kk, v = qkv.float().view(N, S, 3, C).unbind(2)
        s = torch.bmm(q, kk.transpose(1, 2)) * (C ** -0.5)
        lse.copy_(torch.logsumexp(s, dim=-1))
        out.copy_(torch.bmm(torch.softmax(s, -1), v).view(out.shape).to(out.dtype))
        self.launches += 1

    def attn_bwd(self, qkv, out, dout, lse, delta, dqkv, N, S, C):
        x = qkv.float().view(N, S, 3, C).clone().requires_grad_(True)
        with torch.enable_grad():
            q, kk, v = x.unbind(2)
            o = torch.bmm(torch.softmax(torch.bmm(q, kk.transpose(1, 2)) * (C ** -0.5), -1), v)
            o.backward(dout.float().view(N, S, C))
        dqkv.copy_(x.grad.view(dqkv.shape).to(dqkv.dtype))
        self.launches += 3

    def upsample_nearest(self, x, out):
        fy, fx = out.shape[1] // x.shape[1], out.shape[2] // x.shape[2]
        out.copy_(x.repeat_interleave(fy, dim=1).repeat_interleave(fx, dim=2))
        self.launches += 1

    def upsample_nearest_bwd(self, dout, din):
        N, H, W, C = din.shape
        fy, fx = dout.shape[1] // H, dout.shape[2] // W
        din.copy_(dout.float().view(N, H, fy, W, fx, C).sum(dim=(2, 4)).to(din.dtype))
        self.launches += 1

    def image_affine(self, x, out, scale, shift):
        out.copy_(x.float() * scale + shift)
        self.launches += 1

    def mha_fwd(self, qkv, out, lse, N, S, C, heads):
        hd = C // heads
        q, kk, v = (t.reshape(N, S, heads, hd).permute(0, 2, 1, 3) for t in qkv.float().view(N, S, 3, C).unbind(2))
        s = torch.matmul(q, kk.transpose(-1, -2)) * (hd ** -0.5)
        lse.copy_(torch.logsumexp(s, dim=-1))
        out.copy_(torch.matmul(torch.softmax(s, -1), v).permute(0, 2, 1, 3).reshape(out.shape).to(out.dtype))
        self.launches += 1

    def mha_bwd(self, qkv, out, dout, lse, delta, dqkv, N, S, C, heads):
        hd = C // heads
        x = qkv.float().view(N, S, 3, C).clone().requires_grad_(True)
        with torch.enable_grad():
            q, kk, v = (t.reshape(N, S, heads, hd).permute(0, 2, 1, 3) for t in x.unbind(2))
            o = torch.matmul(torch.softmax(torch.matmul(q, kk.transpose(-1, -2)) * (hd ** -0.5), -1), v).permute(0, 2, 1, 3).reshape(N, S, C)
            o.backward(dout.float().view(N, S, C))
        dqkv.copy_(x.grad.view(dqkv.shape).to(dqkv.dtype))
        self.launches += 2

    # ---- GroupNorm family ----
    @staticmethod
    def _cat(x0, x1):
        return (x0 if x1 is None else torch.cat([x0, x1], -1)).float()

    def gn_stats(self, x0, x1, N, HW, G, sums):
        x = self._cat(x0, x1).reshape(N, HW, G, -1).double()
        sums[:, :, 0] = x.sum(dim=(1, 3))
        sums[:, :, 1] = (x * x).sum(dim=(1, 3))
        self.launches += 1

    @staticmethod
    def _gn_formula(x, N, HW, G, sums, gamma, beta, eps, act):
        C = x.shape[-1]
        cnt = (C // G) * HW
        mean = sums[:, :, 0] / cnt
        var = (sums[:, :, 1] / cnt - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        xh = (x.reshape(N, HW, G, C // G) - mean.float()[:, None, :, None]) * rstd.float()[:, None, :, None]
        z = xh.reshape(N, HW, C) * gamma.float() + beta.float()
        return _swish(z) if act else z

    def gn_group_sums(self, cs0, cs1, N, G, sums):
        cs = cs0 if cs1 is None else torch.cat([cs0, cs1], 1)
        sums.copy_(cs.reshape(N, G, -1, 2).sum(2))
        self.launches += 1

    def gn_apply(self, x0, x1, N, HW, G, sums, gamma, beta, eps, act, p_drop, seed, out):
        assert p_drop == 0, "the test double does not model the counter-based dropout stream"
        y = self._gn_formula(self._cat(x0, x1).reshape(N, HW, -1), N, HW, G, sums, gamma, beta, eps, act)
        out.copy_(y.view(out.shape).to(out.dtype))
        self.launches += 1

    def gn_bwd(self, x0, x1, N, HW, G, sums, gamma, beta, eps, act, p_drop, seed, dy, gsums, dgamma, dbeta,
               add, acc0, acc1, dx0, dx1, cs_total=None, cs_per_n=None, cs_n=None, overwrite_dy=False):
        assert p_drop == 0
        x = self._cat(x0, x1).reshape(N, HW, -1).clone().requires_grad_(True)
        gam = gamma.detach().float().clone().requires_grad_(True)
        bet = beta.detach().float().clone().requires_grad_(True)
        C = x.shape[-1]
        with torch.enable_grad():
            xr = x.reshape(N, HW, G, C // G)
            mean = xr.mean(dim=(1, 3), keepdim=True)
            var = xr.var(dim=(1, 3), unbiased=False, keepdim=True)
            z = ((xr - mean) / torch.sqrt(var + eps)).reshape(N, HW, C) * gam + bet
            y = _swish(z) if act else z
            y.backward(dy.float().reshape(N, HW, C))
        dgamma += gam.grad
        dbeta += bet.grad
        dx = x.grad
        if add is not None:
            dx = dx + add.float().reshape(N, HW, C)
        C0 = x0.shape[-1]
        d0 = dx[..., :C0]
        if acc0 is not None:
            d0 = d0 + acc0.float().reshape(N, HW, C0)
        dx0.copy_(d0.reshape(dx0.shape).to(dx0.dtype))
        if x1 is not None:
            d1 = dx[..., C0:]
            if acc1 is not None:
                d1 = d1 + acc1.float().reshape(N, HW, -1)
            dx1.copy_(d1.reshape(dx1.shape).to(dx1.dtype))
        if cs_total is not None or cs_per_n is not None:
            full = d0 if x1 is None else torch.cat([d0, d1], -1)
            cs = full.reshape(N, HW, C).sum(1)[:, :(C if cs_n is None else cs_n)]
            if cs_per_n is not None:
                cs_per_n += cs
            if cs_total is not None:
                cs_total += cs.sum(0)
        self.launches += 2

    def colsum(self, t, N, HW, C, per_n, total, nchw=False):
        s = t.float().sum(dim=(2, 3)) if nchw else t.float().reshape(N, HW, C).sum(1)
        if per_n is not None:
            per_n += s
        if total is not None:
            total += s.sum(0)
        self.launches += 1

    # ---- embedding path ----
    def linear_fwd(self, x, w, b, y, in_swish=False, accumulate=False):
        v = F.linear(_swish(x) if in_swish else x, w, b)
        if accumulate:
            y += v
        else:
            y.copy_(v)
        self.launches += 1

    def linear_bwd_x(self, dy, w, x_pre, dx, accumulate=False):
        v = dy @ w
        if x_pre is not None:
            s = torch.sigmoid(x_pre)
            v = v * (s * (1 + x_pre * (1 - s)))
        if accumulate:
            dx += v
        else:
            dx.copy_(v)
        self.launches += 1

    def linear_bwd_w(self, dy, x, dw, db, in_swish=False):
        xv = _swish(x) if in_swish else x
        dw += (dy.t() @ xv).reshape(dw.shape)
        if db is not None:
            db += dy.sum(0)
        self.launches += 1

    def embedding_fwd(self, table, idx, out):
        out.copy_(table[idx])
        self.launches += 1

    def embedding_bwd(self, dout, idx, dtable, padding_idx=-1):
        m = idx != padding_idx
        dtable.view(-1, dout.shape[1]).index_add_(0, idx[m], dout[m])
        self.launches += 1

    # ---- packing ----
    def gather_pack(self, src, ia, ib, out):
        a = ia.long()
        v = torch.where(a >= 0, src[a.clamp_min(0)], torch.zeros((), dtype=src.dtype))
        if ib is not None:
            b = ib.long()
            v = v + torch.where(b >= 0, src[b.clamp_min(0)], torch.zeros((), dtype=src.dtype))
        out.copy_(v.to(out.dtype))
        self.launches += 1

    def scatter_unpack(self, packed, inv, dst):
        i = inv.long()
        m = i >= 0
        dst[m] += packed[i[m]]
        self.launches += 1

    # ---- diffusion process ----
    def q_sample(self, x0, noise, t, sab, s1ab, xt):
        sh = [x0.shape[0]] + [1] * (x0.dim() - 1)
        xt.copy_(sab[t].view(sh) * x0 + s1ab[t].view(sh) * noise)
        self.launches += 1

    def mse_fwd(self, pred, noise, loss):
        loss.copy_((pred - noise) ** 2)
        self.launches += 1

    def mse_bwd(self, pred, noise, g, dpred):
        dpred.copy_(2 * (pred - noise) * g)
        self.launches += 1

    def sampler_step(self, x, eps_c, eps_u, z, w, coef, step_ptr, clip_last, nan_flag):
        s = int(step_ptr.item())
        c1, c2, sv = coef[s]
        e = eps_c if eps_u is None else (1. + w) * eps_c - w * eps_u
        r = c1 * x - c2 * e
        if s > 0:
            r = r + sv * z
        if torch.isnan(r).any():
            nan_flag.fill_(1)
        if s == 0 and clip_last:
            r = r.clamp(-1, 1)
        x.copy_(r)
        self.launches += 1

    def add_int(self, p, delta):
        p += delta
        self.launches += 1

    # ---- optimizer ----
    def sqnorm(self, g, out, accumulate=False):
        v = (g.double() ** 2).sum().reshape(out.shape)
        if accumulate:
            out += v
        else:
            out.copy_(v)
        self.launches += 1

    def adamw_flat(self, p, g, m, v, sqnorm, max_norm, lr, b1, b2, eps, wd, step):
        if max_norm > 0:
            clip = min(1.0, max_norm / (float(sqnorm.sqrt()) + 1e-6))
            g *= clip
        p *= 1 - lr * wd
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        p.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)
        self.launches += 1
