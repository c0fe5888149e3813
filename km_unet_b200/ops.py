"""torch.autograd bridges to the C ABI (one C call per autograd node).

PyTorch is used for plumbing only: device memory (the caching allocator owns every buffer), the current stream and
autograd bookkeeping.  All arithmetic happens in libkmunet.so.  CPU tensors raise -- there is no fallback.
"""
import ctypes as C
import os

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib
from ._lib import (ScDesc, TnBwdArgs, TnDesc, TnFwdArgs, PwDesc, BnMixBwdArgs, BnMixDesc, BnMixFwdArgs, DwDesc, DagemBn, DagemBwdArgs, DagemDesc, DagemFwdArgs, DeformDesc, DysBwdArgs, DysDesc, DysFwdArgs, HsmBwdArgs, HsmDesc, HsmFwdArgs, KanBwdArgs, KanDesc, KanFwdArgs,
                   KMU_PREC_BF16, KMU_PREC_FP32, check, ptr, stream_ptr)

__all__ = ["kanconv2d", "kanlinear", "layernorm1d", "hsmssd", "dysample", "dysample_sample", "dagem_gate", "bnmix", "dwconv3x3", "pwconv", "pwconv_supported", "triplenorm", "qkv_gate", "smallconv", "smallconv_supported", "iwp", "iwp_supported", "combine3", "combine3_supported", "resize_bilinear_ac", "groupnorm", "groupnorm_supported", "dwconv_bnmix", "lerpmix", "lerpmix_supported", "KMU_PREC_FP32", "KMU_PREC_BF16"]


# ------------------------------------------------------------------------------------------------------ op-level timing
_PROF = None


def profile_start():
    """Record a CUDA-event pair (on the launching stream) around every C-ABI call until profile_stop()."""
    global _PROF
    _PROF = []


def profile_stop():
    """-> {(entry point, shape key): {"calls": n, "ms": total device ms}}; synchronises."""
    global _PROF
    rec, _PROF = _PROF or [], None
    torch.cuda.synchronize()
    out = {}
    for name, key, a, b in rec:
        e = out.setdefault((name, key), {"calls": 0, "ms": 0.0})
        e["calls"] += 1
        e["ms"] += a.elapsed_time(b)
    return out


def _call(name, key, fn, *args):
    if _PROF is None:
        return fn(*args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn(*args)
    b.record()
    _PROF.append((name, key, a, b))
    return r


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------------ KANConv2d
class _KanConv2dFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, base_weight, spline_weight, spline_scaler, grid, ksize, stride, padding, grid_size, spline_order,
                precision, grid_info):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cin, H, W = x.shape
        Cout = base_weight.shape[0]
        uniform, t0, h = grid_info if grid_info is not None else (False, 0.0, 0.0)
        desc = KanDesc(B, Cin, H, W, Cout, ksize, stride, padding, grid_size, spline_order, precision,
                       0 if spline_scaler is None else 1, 1 if uniform else 0, float(t0), float(h))
        Ho = (H + 2 * padding - ksize) // stride + 1
        Wo = (W + 2 * padding - ksize) // stride + 1
        bw, sw, gr = base_weight.contiguous(), spline_weight.contiguous(), grid.contiguous()
        sc = None if spline_scaler is None else spline_scaler.contiguous()
        nbytes = lib.kmu_kanconv2d_fwd_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            raise RuntimeError("kanconv2d: " + _lib.last_error())
        ws = _workspace(nbytes, x.device)
        y = torch.empty(B, Cout, Ho, Wo, dtype=torch.float32, device=x.device)
        args = KanFwdArgs(desc, ptr(x), ptr(bw), ptr(sw), ptr(sc), ptr(gr), ptr(y), ws.data_ptr(), ws.numel())
        check(_call("kmu_kanconv2d_fwd", (B, Cin, H, W, Cout), lib.kmu_kanconv2d_fwd, C.byref(args), stream_ptr()), "kmu_kanconv2d_fwd")
        ctx.save_for_backward(x, bw, sw, sc, gr)
        ctx.desc = desc
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, bw, sw, sc, gr = ctx.saved_tensors
        desc = ctx.desc
        dy = dy.to(torch.float32).contiguous()
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[1:4])
        dx = torch.empty_like(x) if need_x else None
        dbw = torch.empty_like(bw) if need_w else None
        dsw = torch.empty_like(sw) if need_w else None
        dsc = torch.empty_like(sc) if (need_w and sc is not None) else None
        nbytes = lib.kmu_kanconv2d_bwd_workspace_bytes(C.byref(desc))
        ws = _workspace(nbytes, x.device)
        args = KanBwdArgs(desc, ptr(x), ptr(dy), ptr(bw), ptr(sw), ptr(sc), ptr(gr), ptr(dx), ptr(dbw), ptr(dsw), ptr(dsc),
                          ws.data_ptr(), ws.numel())
        check(_call("kmu_kanconv2d_bwd", (desc.B, desc.Cin, desc.H, desc.W, desc.Cout), lib.kmu_kanconv2d_bwd, C.byref(args), stream_ptr()),
              "kmu_kanconv2d_bwd")
        return dx, dbw, dsw, dsc, None, None, None, None, None, None, None, None


def grid_info(grid):
    """(uniform_and_shared, t0, h) of a KANLinear knot table (in_features, n_knots): True when every row is the same
    uniform knot vector.  One small D2H copy; callers cache it per grid version."""
    g = grid.detach().double().cpu()
    n = g.shape[1]
    t0 = g[0, 0].item()
    h = ((g[0, -1] - g[0, 0]) / (n - 1)).item()
    if not h > 0:
        return (False, t0, h)
    ideal = t0 + h * torch.arange(n, dtype=torch.float64)
    return (bool((g - ideal).abs().max().item() <= 1e-5 * h), t0, h)


def kanconv2d(x, base_weight, spline_weight, spline_scaler, grid, kernel_size, stride=1, padding=0, grid_size=5,
              spline_order=3, precision=KMU_PREC_FP32, grid_meta=None):
    """KANConv2d.forward (convKAN/KANConv2Dlayers.py:15-37): x (B,Cin,H,W) -> (B,Cout,Ho,Wo).
    grid_meta = grid_info(grid) lets the bf16 precision take the tcgen05 path; None keeps the fp32 family."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.kanconv2d: CUDA tensors only (no CPU fallback)")
    return _KanConv2dFn.apply(x, base_weight, spline_weight, spline_scaler, grid, int(kernel_size), int(stride), int(padding),
                              int(grid_size), int(spline_order), int(precision), grid_meta)


def kanlinear(x, base_weight, spline_weight, spline_scaler, grid, grid_size=5, spline_order=3, precision=KMU_PREC_FP32,
              grid_meta=None):
    """KANLinear.forward (convKAN/KANlayers.py:652-660): x (M,in) -> (M,out), as a 1x1 'convolution' over M pixels."""
    M, F = x.shape
    y = kanconv2d(x.reshape(M, F, 1, 1), base_weight, spline_weight, spline_scaler, grid, 1, 1, 0, grid_size, spline_order,
                  precision, grid_meta)
    return y.reshape(M, -1)


# ------------------------------------------------------------------------------------------------------ LayerNorm1D
class _LayerNorm1dFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, eps):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, L = x.shape
        w, b = weight.reshape(-1).contiguous(), bias.reshape(-1).contiguous()
        y = torch.empty_like(x)
        check(_call("kmu_layernorm1d_fwd", (B, Cc, L), lib.kmu_layernorm1d_fwd, ptr(x), ptr(w), ptr(b), ptr(y), None, B, Cc, L, float(eps),
                    stream_ptr()), "kmu_layernorm1d_fwd")
        ctx.save_for_backward(x, w)
        ctx.eps = float(eps)
        ctx.wshape = weight.shape
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, w = ctx.saved_tensors
        B, Cc, L = x.shape
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        db = torch.empty_like(w)
        ws = _workspace(lib.kmu_layernorm1d_bwd_workspace_bytes(B, Cc, L), x.device)
        check(_call("kmu_layernorm1d_bwd", (B, Cc, L), lib.kmu_layernorm1d_bwd, ptr(x), ptr(w), ptr(dy), ptr(dx), ptr(dw), ptr(db), B, Cc, L,
                    ctx.eps, ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_layernorm1d_bwd")
        return dx, dw.reshape(ctx.wshape), db.reshape(ctx.wshape), None


def layernorm1d(x, weight, bias, eps=1e-5):
    """LayerNorm1D.forward (vim_block_init/vim_utils_init.py:50-59) on (B,C,L)."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.layernorm1d: CUDA tensors only (no CPU fallback)")
    return _LayerNorm1dFn.apply(x, weight, bias, eps)


# ------------------------------------------------------------------------------------------------------ HSMSSD
class _HsmssdFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, w_bcdt, w_dw, w_hz, w_out, A, D, state_dim, precision):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, L = x.shape
        H = int(round(L ** 0.5))
        N = int(state_dim)
        desc = HsmDesc(B, Cc, L, H, N, int(precision))
        nbytes = lib.kmu_hsmssd_fwd_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            raise RuntimeError("hsmssd: " + _lib.last_error())
        dev = x.device
        ws = _workspace(nbytes, dev)
        wp, wd, whz, wo = w_bcdt.contiguous(), w_dw.contiguous(), w_hz.contiguous(), w_out.contiguous()
        Ac, Dc = A.contiguous(), D.contiguous()
        y = torch.empty(B, Cc, L, dtype=torch.float32, device=dev)
        h = torch.empty(B, Cc, N, dtype=torch.float32, device=dev)
        # the fused bf16 sweeps (csrc/hsm_fused.cu) keep P = dw3x3(Wp x) in TMEM: nothing to save (402 MB per call at (32,16,16384))
        fused = int(precision) == KMU_PREC_BF16 and os.environ.get("KMU_HSM_FUSED", "1") != "0"
        P = None if fused else torch.empty(B, 3 * N, L, dtype=torch.float32, device=dev)
        stats = torch.empty(B, 2, N, dtype=torch.float32, device=dev)
        hs = torch.empty(B, Cc, N, dtype=torch.float32, device=dev)
        hz = torch.empty(B, 2 * Cc, N, dtype=torch.float32, device=dev)
        args = HsmFwdArgs(desc, ptr(x), ptr(wp), ptr(wd), ptr(whz), ptr(wo), ptr(Ac), ptr(Dc), ptr(y), ptr(h), ptr(P),
                          ptr(stats), ptr(hs), ptr(hz), ws.data_ptr(), ws.numel())
        check(_call("kmu_hsmssd_fwd", (B, Cc, L), lib.kmu_hsmssd_fwd, C.byref(args), stream_ptr()), "kmu_hsmssd_fwd")
        ctx.save_for_backward(x, wp, wd, whz, wo, Ac, Dc, P, stats, hs, hz, h)
        ctx.desc = desc
        ctx.shapes = (w_bcdt.shape, w_dw.shape, w_hz.shape, w_out.shape)
        return y.view(B, Cc, H, H), h

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy, dh):
        lib = _lib.lib()
        x, wp, wd, whz, wo, Ac, Dc, P, stats, hs, hz, h = ctx.saved_tensors
        desc = ctx.desc
        dev = x.device
        dy = dy.to(torch.float32).reshape(x.shape).contiguous()
        dh = None if dh is None else dh.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        dwp, dwd, dwhz, dwo = (torch.empty_like(t) for t in (wp, wd, whz, wo))
        dA, dD = torch.empty_like(Ac), torch.empty_like(Dc)
        nbytes = lib.kmu_hsmssd_bwd_workspace_bytes(C.byref(desc))
        ws = _workspace(nbytes, dev)
        args = HsmBwdArgs(desc, ptr(x), ptr(dy), ptr(dh), ptr(wp), ptr(wd), ptr(whz), ptr(wo), ptr(Ac), ptr(Dc), ptr(P),
                          ptr(stats), ptr(hs), ptr(hz), ptr(h), ptr(dx), ptr(dwp), ptr(dwd), ptr(dwhz), ptr(dwo), ptr(dA),
                          ptr(dD), ws.data_ptr(), ws.numel())
        check(_call("kmu_hsmssd_bwd", (desc.B, desc.C, desc.L), lib.kmu_hsmssd_bwd, C.byref(args), stream_ptr()), "kmu_hsmssd_bwd")
        s = ctx.shapes
        return dx, dwp.reshape(s[0]), dwd.reshape(s[1]), dwhz.reshape(s[2]), dwo.reshape(s[3]), dA, dD, None, None


def hsmssd(x, w_bcdt, w_dw, w_hz, w_out, A, D, state_dim=64, precision=KMU_PREC_FP32):
    """HSMSSD.forward (vim_block_init/efficient_vim_init.py:33-61): x (B,C,L) -> (y (B,C,H,H), h (B,C,N)).
    precision=KMU_PREC_BF16 runs the BCdt projection + depthwise conv of the forward as one tcgen05 convolution."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.hsmssd: CUDA tensors only (no CPU fallback)")
    return _HsmssdFn.apply(x, w_bcdt, w_dw, w_hz, w_out, A, D, state_dim, precision)


# ------------------------------------------------------------------------------------------------------ DySample
class _DySampleFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, w_offset, b_offset, init_pos, scale, groups):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        noff = 2 * groups * scale * scale
        desc = DysDesc(B, Cc, H, W, scale, groups)
        w = w_offset.reshape(noff, Cc).contiguous()
        b = b_offset.contiguous()
        ip = init_pos.reshape(-1).contiguous()
        offset = torch.empty(B, noff, H, W, dtype=torch.float32, device=x.device)
        out = torch.empty(B, Cc, scale * H, scale * W, dtype=torch.float32, device=x.device)
        args = DysFwdArgs(desc, ptr(x), ptr(w), ptr(b), ptr(ip), ptr(offset), ptr(out))
        check(_call("kmu_dysample_fwd", (B, Cc, H, W), lib.kmu_dysample_fwd, C.byref(args), stream_ptr()), "kmu_dysample_fwd")
        ctx.save_for_backward(x, w, offset)
        ctx.desc = desc
        ctx.wshape = w_offset.shape
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        lib = _lib.lib()
        x, w, offset = ctx.saved_tensors
        desc = ctx.desc
        dout = dout.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        db = torch.empty(w.shape[0], dtype=torch.float32, device=x.device)
        ws = _workspace(lib.kmu_dysample_bwd_workspace_bytes(C.byref(desc)), x.device)
        args = DysBwdArgs(desc, ptr(x), ptr(w), ptr(offset), ptr(dout), ptr(dx), ptr(dw), ptr(db), ws.data_ptr(), ws.numel())
        check(_call("kmu_dysample_bwd", (desc.B, desc.C, desc.H, desc.W), lib.kmu_dysample_bwd, C.byref(args), stream_ptr()), "kmu_dysample_bwd")
        return dx, dw.reshape(ctx.wshape), db, None, None, None


def dysample(x, w_offset, b_offset, init_pos, scale=2, groups=4):
    """DySample.forward_lp without dyscope (DySample_md.py:63-68): x (B,C,H,W) -> (B,C,scale*H,scale*W)."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.dysample: CUDA tensors only (no CPU fallback)")
    return _DySampleFn.apply(x, w_offset, b_offset, init_pos, int(scale), int(groups))


class _DySampleSampleFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, offset, scale, groups):
        lib = _lib.lib()
        x, offset = x.contiguous(), offset.contiguous()
        B, Cc, H, W = x.shape
        desc = DysDesc(B, Cc, H, W, scale, groups)
        out = torch.empty(B, Cc, scale * H, scale * W, dtype=torch.float32, device=x.device)
        check(lib.kmu_dysample_sample_fwd(C.byref(desc), ptr(x), ptr(offset), ptr(out), stream_ptr()),
              "kmu_dysample_sample_fwd")
        ctx.save_for_backward(x, offset)
        ctx.desc = desc
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        lib = _lib.lib()
        x, offset = ctx.saved_tensors
        dout = dout.to(torch.float32).contiguous()
        dx = torch.zeros_like(x)
        doff = torch.empty_like(offset)
        check(lib.kmu_dysample_sample_bwd(C.byref(ctx.desc), ptr(x), ptr(offset), ptr(dout), ptr(dx), ptr(doff), stream_ptr()),
              "kmu_dysample_sample_bwd")
        return dx, doff, None, None


def dysample_sample(x, offset, scale=2, groups=4):
    """DySample.sample (DySample_md.py:49-61) for a caller-built offset tensor (B, 2*groups*scale^2, H, W)."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.dysample_sample: CUDA tensors only (no CPU fallback)")
    return _DySampleSampleFn.apply(x, offset, int(scale), int(groups))


# ------------------------------------------------------------------------------------------------------ DAGEM
class _DagemGateFn(torch.autograd.Function):
    """inputs: x, deformed, 9 Linear/conv tensors, 5 x (bn weight, bn bias); non-differentiable: running stats, flags."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, deformed, ea_w, ea_b, vu_w, vu_b, eu_w, eu_b, er_w, er_b, wf, bn_w0, bn_b0, bn_w1, bn_b1, bn_w2, bn_b2,
                bn_w3, bn_b3, bn_w4, bn_b4, running, training, momentum, eps):
        lib = _lib.lib()
        x, deformed = x.contiguous(), deformed.contiguous()
        B, Cc, H, W = x.shape
        desc = DagemDesc(B, Cc, H, W, 1 if training else 0, float(momentum), float(eps))
        nsaved = lib.kmu_dagem_saved_bytes(C.byref(desc))
        if nsaved == 0:
            raise RuntimeError("dagem: " + _lib.last_error())
        lin = [t.contiguous() for t in (ea_w, ea_b, vu_w, vu_b, eu_w, eu_b, er_w, er_b, wf)]
        bnw = [t.contiguous() for t in (bn_w0, bn_w1, bn_w2, bn_w3, bn_w4)]
        bnb = [t.contiguous() for t in (bn_b0, bn_b1, bn_b2, bn_b3, bn_b4)]
        saved = torch.empty(nsaved // 4, dtype=torch.float32, device=x.device)
        out = torch.empty_like(x)
        ws = _workspace(lib.kmu_dagem_fwd_workspace_bytes(C.byref(desc)), x.device)
        bns = (DagemBn * 5)()
        for i in range(5):
            rm, rv = running[i]
            bns[i] = DagemBn(ptr(bnw[i]), ptr(bnb[i]), ptr(rm), ptr(rv))
        args = DagemFwdArgs(desc, ptr(x), ptr(deformed), *[ptr(t) for t in lin], bns, ptr(out), ptr(saved), ws.data_ptr(),
                            ws.numel())
        check(_call("kmu_dagem_fwd", (B, Cc, H, W), lib.kmu_dagem_fwd, C.byref(args), stream_ptr()), "kmu_dagem_fwd")
        ctx.save_for_backward(x, deformed, saved, *lin, *bnw)
        ctx.desc = desc
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        lib = _lib.lib()
        x, deformed, saved = ctx.saved_tensors[:3]
        ea_w, ea_b, vu_w, vu_b, eu_w, eu_b, er_w, er_b, wf = ctx.saved_tensors[3:12]
        bnw = ctx.saved_tensors[12:17]
        desc = ctx.desc
        dout = dout.to(torch.float32).contiguous()
        dx, dd = torch.empty_like(x), torch.empty_like(x)
        dlin = [torch.empty_like(t) for t in (ea_w, ea_b, vu_w, vu_b, eu_w, eu_b, er_w, er_b, wf)]
        dbw = [torch.empty_like(t) for t in bnw]
        dbb = [torch.empty_like(t) for t in bnw]
        ws = _workspace(lib.kmu_dagem_bwd_workspace_bytes(C.byref(desc)), x.device)
        d_ea_w, d_ea_b, d_vu_w, d_vu_b, d_eu_w, d_eu_b, d_er_w, d_er_b, d_wf = dlin
        args = DagemBwdArgs(desc, ptr(x), ptr(deformed), ptr(dout), ptr(saved), ptr(ea_w), ptr(vu_w), ptr(eu_w), ptr(er_w), ptr(wf),
                            ptr(dx), ptr(dd), ptr(d_ea_w), ptr(d_ea_b), ptr(d_vu_w), ptr(d_vu_b), ptr(d_eu_w), ptr(d_eu_b),
                            ptr(d_er_w), ptr(d_er_b), ptr(d_wf), (_lib._f32p * 5)(*[ptr(t) for t in dbw]),
                            (_lib._f32p * 5)(*[ptr(t) for t in dbb]), ws.data_ptr(), ws.numel())
        check(_call("kmu_dagem_bwd", (desc.B, desc.C, desc.H, desc.W), lib.kmu_dagem_bwd, C.byref(args), stream_ptr()), "kmu_dagem_bwd")
        bn_grads = []
        for i in range(5):
            bn_grads += [dbw[i], dbb[i]]
        return (dx, dd, *dlin, *bn_grads, None, None, None, None)


def dagem_gate(x, deformed, linears, bns, training, momentum=0.1, eps=1e-5):
    """The gating + final-aggregation part of DAGEM.forward (DAGEM_md.py:62-92,104-110).
    linears = (ea_w, ea_b, vu_w, vu_b, eu_w, eu_b, er_w, er_b, wf); bns = 5 x (weight, bias, running_mean, running_var) in
    the order edge_aggregation, edge_update, vertex_update, update_edge_reduce, final_aggregation."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.dagem_gate: CUDA tensors only (no CPU fallback)")
    flat = []
    for w, b, _, _ in bns:
        flat += [w, b]
    running = [(rm, rv) for _, _, rm, rv in bns]
    return _DagemGateFn.apply(x, deformed, *linears, *flat, running, bool(training), momentum, eps)


class _DeformConvFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, offset, weight, bias):
        lib = _lib.lib()
        x, offset, weight = x.contiguous(), offset.contiguous(), weight.contiguous()
        B, Cc, H, W = x.shape
        desc = DeformDesc(B, Cc, H, W, weight.shape[0])
        out = torch.empty(B, weight.shape[0], H, W, dtype=torch.float32, device=x.device)
        check(_call("kmu_deformconv3x3_fwd", (B, Cc, H, W), lib.kmu_deformconv3x3_fwd, C.byref(desc), ptr(x), ptr(offset), ptr(weight),
                    ptr(bias.contiguous()) if bias is not None else None, ptr(out), stream_ptr()), "kmu_deformconv3x3_fwd")
        ctx.save_for_backward(x, offset, weight)
        ctx.desc, ctx.has_bias = desc, bias is not None
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        lib = _lib.lib()
        x, offset, weight = ctx.saved_tensors
        desc = ctx.desc
        dout = dout.to(torch.float32).contiguous()
        dx, doff, dw = torch.empty_like(x), torch.empty_like(offset), torch.empty_like(weight)
        db = torch.empty(weight.shape[0], dtype=torch.float32, device=x.device) if ctx.has_bias else None
        ws = _workspace(lib.kmu_deformconv3x3_bwd_workspace_bytes(C.byref(desc)), x.device)
        check(_call("kmu_deformconv3x3_bwd", (desc.B, desc.C, desc.H, desc.W), lib.kmu_deformconv3x3_bwd, C.byref(desc), ptr(x), ptr(offset),
                    ptr(weight), ptr(dout), ptr(dx), ptr(doff), ptr(dw), ptr(db) if db is not None else None, ws.data_ptr(), ws.numel(),
                    stream_ptr()), "kmu_deformconv3x3_bwd")
        return dx, doff, dw, db


def deformconv3x3_supported(x, weight):
    return (x.is_cuda and x.dim() == 4 and x.shape[2] * x.shape[3] <= 4096 and tuple(weight.shape[2:]) == (3, 3)
            and x.shape[1] <= 256 and weight.shape[0] <= 256 and weight.shape[1] == x.shape[1])


def deformconv3x3(x, offset, weight, bias=None):
    """torchvision.ops.deform_conv2d(x, offset, weight, bias, padding=1) for a 3x3 kernel, stride 1, one offset group
    (DAGEM_md.py:46,98-101) on the caller's stream, deterministic backward."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.deformconv3x3: CUDA tensors only (no CPU fallback)")
    return _DeformConvFn.apply(x, offset, weight, bias)


# ------------------------------------------------------------------------------------------------------ EfficientViMBlock shell
class _BnMixFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, res, alpha, running_mean, running_var, training, momentum, eps, relu):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc = x.shape[0], x.shape[1]
        HW = x.numel() // (B * Cc)
        mix = res is not None
        desc = BnMixDesc(B, Cc, HW, 1 if training else 0, 1 if relu else 0, 1 if mix else 0, float(momentum), float(eps))
        w, b = weight.contiguous(), bias.contiguous()
        r = res.contiguous() if mix else None
        al = alpha.contiguous() if mix else None
        y = torch.empty_like(x)
        stat = torch.empty(Cc, 2, dtype=torch.float32, device=x.device)
        ws = _workspace(lib.kmu_bnmix_workspace_bytes(C.byref(desc)), x.device)
        args = BnMixFwdArgs(desc, ptr(x), ptr(w), ptr(b), ptr(running_mean), ptr(running_var), ptr(r), ptr(al), ptr(y), ptr(stat),
                            ws.data_ptr(), ws.numel())
        check(_call("kmu_bnmix_fwd", (B, Cc, HW), lib.kmu_bnmix_fwd, C.byref(args), stream_ptr()), "kmu_bnmix_fwd")
        ctx.save_for_backward(x, w, b, stat, r, al)
        ctx.desc = desc
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, w, b, stat, r, al = ctx.saved_tensors
        desc = ctx.desc
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        dw, db = torch.empty_like(w), torch.empty_like(b)
        dres = torch.empty_like(x) if desc.mix else None
        dal = torch.empty_like(al) if desc.mix else None
        ws = _workspace(lib.kmu_bnmix_workspace_bytes(C.byref(desc)), x.device)
        args = BnMixBwdArgs(desc, ptr(x), ptr(dy), ptr(w), ptr(b), ptr(stat), ptr(r), ptr(al), ptr(dx), ptr(dw), ptr(db), ptr(dres),
                            ptr(dal), ws.data_ptr(), ws.numel())
        check(_call("kmu_bnmix_bwd", (desc.B, desc.C, desc.HW), lib.kmu_bnmix_bwd, C.byref(args), stream_ptr()), "kmu_bnmix_bwd")
        return dx, dw, db, dres, dal, None, None, None, None, None, None


def bnmix(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5, relu=False, res=None, alpha=None):
    """BatchNorm2d (batch statistics when `training`, running-stat update in place) with the fused epilogues of the
    EfficientViMBlock shell: ReLU (ffn.fc1) or the layer-scale mix (1 - sigmoid(alpha)) res + sigmoid(alpha) BN(x)
    (vim_block_init/efficient_vim_init.py:85,93,96).  alpha is the raw (pre-sigmoid) (C,) row."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.bnmix: CUDA tensors only (no CPU fallback)")
    if (res is None) != (alpha is None):
        raise ValueError("bnmix: res and alpha go together")
    return _BnMixFn.apply(x, weight, bias, res, alpha, running_mean, running_var, bool(training), momentum, eps, bool(relu))


class _DwBnMixFn(torch.autograd.Function):
    """y = (1 - sigmoid(alpha)) x + sigmoid(alpha) BN(dwconv3x3(x)): EfficientViMBlock's depthwise branches
    (vim_block_init/efficient_vim_init.py:85,93) as ONE autograd node, so that the mix's d(res) is added to the convolution's input
    gradient inside the dx kernel (kmu_dwconv3x3_bwd_add) instead of by an autograd accumulation kernel."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, conv_w, weight, bias, alpha, running_mean, running_var, training, momentum, eps):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        HW = H * W
        ddesc = DwDesc(B, Cc, H, W)
        cw = conv_w.reshape(Cc, 9).contiguous()
        t = torch.empty_like(x)
        check(_call("kmu_dwconv3x3_fwd", (B, Cc, H, W), lib.kmu_dwconv3x3_fwd, C.byref(ddesc), ptr(x), ptr(cw), None, ptr(t), stream_ptr()),
              "kmu_dwconv3x3_fwd")
        desc = BnMixDesc(B, Cc, HW, 1 if training else 0, 0, 1, float(momentum), float(eps))
        w, b, al = weight.contiguous(), bias.contiguous(), alpha.contiguous()
        y = torch.empty_like(x)
        stat = torch.empty(Cc, 2, dtype=torch.float32, device=x.device)
        ws = _workspace(lib.kmu_bnmix_workspace_bytes(C.byref(desc)), x.device)
        args = BnMixFwdArgs(desc, ptr(t), ptr(w), ptr(b), ptr(running_mean), ptr(running_var), ptr(x), ptr(al), ptr(y), ptr(stat),
                            ws.data_ptr(), ws.numel())
        check(_call("kmu_bnmix_fwd", (B, Cc, HW), lib.kmu_bnmix_fwd, C.byref(args), stream_ptr()), "kmu_bnmix_fwd")
        ctx.save_for_backward(x, cw, t, w, b, stat, al)
        ctx.desc, ctx.ddesc, ctx.cwshape = desc, ddesc, conv_w.shape
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, cw, t, w, b, stat, al = ctx.saved_tensors
        desc, ddesc = ctx.desc, ctx.ddesc
        dy = dy.to(torch.float32).contiguous()
        dt = torch.empty_like(x)
        dres = torch.empty_like(x)
        dw, db, dal = torch.empty_like(w), torch.empty_like(b), torch.empty_like(al)
        ws = _workspace(lib.kmu_bnmix_workspace_bytes(C.byref(desc)), x.device)
        args = BnMixBwdArgs(desc, ptr(t), ptr(dy), ptr(w), ptr(b), ptr(stat), ptr(x), ptr(al), ptr(dt), ptr(dw), ptr(db), ptr(dres),
                            ptr(dal), ws.data_ptr(), ws.numel())
        check(_call("kmu_bnmix_bwd", (desc.B, desc.C, desc.HW), lib.kmu_bnmix_bwd, C.byref(args), stream_ptr()), "kmu_bnmix_bwd")
        dcw = torch.empty_like(cw)
        ws2 = _workspace(lib.kmu_dwconv3x3_bwd_workspace_bytes(C.byref(ddesc)), x.device)
        # dx = conv^T(dt) + dres, written over dres
        check(_call("kmu_dwconv3x3_bwd", (ddesc.B, ddesc.C, ddesc.H, ddesc.W), lib.kmu_dwconv3x3_bwd_add, C.byref(ddesc), ptr(x), ptr(dt),
                    ptr(cw), ptr(dres), ptr(dres), ptr(dcw), None, ws2.data_ptr(), ws2.numel(), stream_ptr()), "kmu_dwconv3x3_bwd_add")
        return dres, dcw.reshape(ctx.cwshape), dw, db, dal, None, None, None, None, None


def dwconv_bnmix(x, conv_weight, bn_weight, bn_bias, alpha, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    """(1 - sigmoid(alpha)) x + sigmoid(alpha) BatchNorm2d(dwconv3x3(x)) with x feeding both the convolution and the mix."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.dwconv_bnmix: CUDA tensors only (no CPU fallback)")
    return _DwBnMixFn.apply(x, conv_weight, bn_weight, bn_bias, alpha, running_mean, running_var, bool(training), momentum, eps)


class _DwConv3x3Fn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        desc = DwDesc(B, Cc, H, W)
        w = weight.reshape(Cc, 9).contiguous()
        b = None if bias is None else bias.contiguous()
        y = torch.empty_like(x)
        check(_call("kmu_dwconv3x3_fwd", (B, Cc, H, W), lib.kmu_dwconv3x3_fwd, C.byref(desc), ptr(x), ptr(w), ptr(b), ptr(y),
                    stream_ptr()), "kmu_dwconv3x3_fwd")
        ctx.save_for_backward(x, w)
        ctx.desc, ctx.has_bias, ctx.wshape = desc, bias is not None, weight.shape
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, w = ctx.saved_tensors
        desc = ctx.desc
        dy = dy.to(torch.float32).contiguous()
        need_x = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty(desc.C, dtype=torch.float32, device=x.device) if (need_w and ctx.has_bias) else None
        ws = _workspace(lib.kmu_dwconv3x3_bwd_workspace_bytes(C.byref(desc)), x.device)
        check(_call("kmu_dwconv3x3_bwd", (desc.B, desc.C, desc.H, desc.W), lib.kmu_dwconv3x3_bwd, C.byref(desc), ptr(x), ptr(dy), ptr(w),
                    ptr(dx), ptr(dw), ptr(db), ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_dwconv3x3_bwd")
        return dx, None if dw is None else dw.reshape(ctx.wshape), db


class _DwConv3x3ScaledFn(torch.autograd.Function):
    """y = scale[b, c] * (dwconv3x3(x) + bias): the convolution of DirectionAttention with its squeeze-excite factor folded in."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, scale):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        desc = DwDesc(B, Cc, H, W)
        w = weight.reshape(Cc, 9).contiguous()
        b = None if bias is None else bias.contiguous()
        sc = scale.reshape(B, Cc).contiguous()
        y = torch.empty_like(x)
        check(_call("kmu_dwconv3x3_fwd", (B, Cc, H, W), lib.kmu_dwconv3x3_scaled_fwd, C.byref(desc), ptr(x), ptr(w), ptr(b), ptr(sc),
                    ptr(y), stream_ptr()), "kmu_dwconv3x3_scaled_fwd")
        ctx.save_for_backward(x, w, sc, b if b is not None else w.new_zeros(0))
        ctx.desc, ctx.has_bias, ctx.wshape, ctx.sshape = desc, bias is not None, weight.shape, scale.shape
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, w, sc, b = ctx.saved_tensors
        desc = ctx.desc
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty(desc.C, dtype=torch.float32, device=x.device) if ctx.has_bias else None
        dsc = torch.empty_like(sc)
        ws = _workspace(lib.kmu_dwconv3x3_bwd_workspace_bytes(C.byref(desc)), x.device)
        check(_call("kmu_dwconv3x3_bwd", (desc.B, desc.C, desc.H, desc.W), lib.kmu_dwconv3x3_scaled_bwd, C.byref(desc), ptr(x), ptr(dy),
                    ptr(w), ptr(b) if ctx.has_bias else None, ptr(sc), ptr(dx), ptr(dw), ptr(db), ptr(dsc), ws.data_ptr(), ws.numel(),
                    stream_ptr()), "kmu_dwconv3x3_scaled_bwd")
        return dx, dw.reshape(ctx.wshape), db, dsc.reshape(ctx.sshape)


def dwconv3x3(x, weight, bias=None, scale=None):
    """Depthwise 3x3 convolution, stride 1, zero padding 1: weight (C,1,3,3), optional bias (C).  scale (B,C) multiplies the
    result per plane (y = scale * (conv + bias)) inside the same kernels."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.dwconv3x3: CUDA tensors only (no CPU fallback)")
    if scale is not None:
        return _DwConv3x3ScaledFn.apply(x, weight, bias, scale)
    return _DwConv3x3Fn.apply(x, weight, bias)


class _PwConvFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, precision):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cin = x.shape[0], x.shape[1]
        HW = x.numel() // (B * Cin)
        Cout = weight.shape[0]
        desc = PwDesc(B, Cin, Cout, HW)
        w = weight.reshape(Cout, Cin).contiguous()
        b = None if bias is None else bias.contiguous()
        y = torch.empty((B, Cout) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
        from . import config
        wide_tc = config.conv_wide == "tc"
        tc = (precision == KMU_PREC_BF16 or (config.conv_fwd == "tma" and wide_tc)) and bool(lib.kmu_pwconv_tc_supported(C.byref(desc)))
        if config.conv_fwd == "tma" and x.data_ptr() % 16 == 0 and bool(lib.kmu_pwconv_tma_fwd_supported(C.byref(desc))):
            tc = False
            ws = _workspace(lib.kmu_pwconv_tma_fwd_workspace_bytes(C.byref(desc)), x.device)
            check(_call("kmu_pwconv_tma_fwd", (B, Cin, Cout, HW), lib.kmu_pwconv_tma_fwd, C.byref(desc), ptr(x), ptr(w), ptr(b), ptr(y),
                        ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_pwconv_tma_fwd")
        elif tc:
            ws = _workspace(lib.kmu_pwconv_tc_workspace_bytes(C.byref(desc)), x.device)
            check(_call("kmu_pwconv_tc_fwd", (B, Cin, Cout, HW), lib.kmu_pwconv_tc_fwd, C.byref(desc), ptr(x), ptr(w), ptr(b), ptr(y),
                        ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_pwconv_tc_fwd")
        else:
            check(_call("kmu_pwconv_fwd", (B, Cin, Cout, HW), lib.kmu_pwconv_fwd, C.byref(desc), ptr(x), ptr(w), ptr(b), ptr(y),
                        stream_ptr()), "kmu_pwconv_fwd")
        ctx.save_for_backward(x, w)
        ctx.desc, ctx.has_bias, ctx.wshape = desc, bias is not None, weight.shape
        ctx.tc = tc or (config.conv_bwd == "fused" and wide_tc and bool(lib.kmu_pwconv_tc_supported(C.byref(desc))))
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, w = ctx.saved_tensors
        desc = ctx.desc
        dy = dy.to(torch.float32).contiguous()
        need_x = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty(desc.Cout, dtype=torch.float32, device=x.device) if (need_w and ctx.has_bias) else None
        key = (desc.B, desc.Cin, desc.Cout, desc.HW)
        from . import config
        if (config.conv_bwd == "fused" and need_x and need_w and dy.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0
                and bool(lib.kmu_pwconv_fused_bwd_supported(C.byref(desc)))):
            ws = _workspace(lib.kmu_pwconv_fused_bwd_workspace_bytes(C.byref(desc)), x.device)
            check(_call("kmu_pwconv_fused_bwd", key, lib.kmu_pwconv_fused_bwd, C.byref(desc), ptr(x), ptr(dy), ptr(w), ptr(dx),
                        ptr(dw), ptr(db), ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_pwconv_fused_bwd")
        elif ctx.tc:
            tc_w = bool(lib.kmu_pwconv_tc_wgrad_supported(C.byref(desc)))
            ws = _workspace(lib.kmu_pwconv_tc_workspace_bytes(C.byref(desc)), x.device)
            check(_call("kmu_pwconv_tc_bwd", key, lib.kmu_pwconv_tc_bwd, C.byref(desc), ptr(x), ptr(dy), ptr(w), ptr(dx),
                        ptr(dw) if tc_w else None, ptr(db) if tc_w else None, ws.data_ptr(), ws.numel(), stream_ptr()),
                  "kmu_pwconv_tc_bwd")
            if need_w and not tc_w:          # Cin > 240: the weight gradient stays on the fp32 kernel
                ws2 = _workspace(lib.kmu_pwconv_bwd_workspace_bytes(C.byref(desc)), x.device)
                check(_call("kmu_pwconv_bwd", key, lib.kmu_pwconv_bwd, C.byref(desc), ptr(x), ptr(dy), ptr(w), None, ptr(dw), ptr(db),
                            ws2.data_ptr(), ws2.numel(), stream_ptr()), "kmu_pwconv_bwd")
        else:
            ws = _workspace(lib.kmu_pwconv_bwd_workspace_bytes(C.byref(desc)), x.device)
            check(_call("kmu_pwconv_bwd", key, lib.kmu_pwconv_bwd, C.byref(desc), ptr(x), ptr(dy), ptr(w), ptr(dx), ptr(dw), ptr(db),
                        ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_pwconv_bwd")
        return dx, None if dw is None else dw.reshape(ctx.wshape), db, None


def pwconv_supported(cin, cout):
    """True when the CUDA weight-gradient kernel covers this channel pair (forward / input gradient always do)."""
    d = PwDesc(1, int(cin), int(cout), 1)
    return bool(_lib.lib().kmu_pwconv_wgrad_supported(C.byref(d)))


def pwconv(x, weight, bias=None, precision=KMU_PREC_FP32):
    """1x1 convolution on NCHW: weight (Cout,Cin,1,1) or (Cout,Cin), optional bias (Cout).  precision=KMU_PREC_BF16 takes the
    tcgen05 kernels where the channel counts allow (multiples of 16 up to 256), the fp32 kernels otherwise."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.pwconv: CUDA tensors only (no CPU fallback)")
    return _PwConvFn.apply(x, weight, bias, int(precision))


# ------------------------------------------------------------------------------------------------------ caller-side fusions
class _TripleNormFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, gh, bh, gw, bw, gc, bc, eps_gn, eps_ln):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc = x.shape[0], x.shape[1]
        HW = x.numel() // (B * Cc)
        desc = TnDesc(B, Cc, HW, float(eps_gn), float(eps_ln))
        nbytes = lib.kmu_triplenorm_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            raise RuntimeError("triplenorm: " + _lib.last_error())
        ps = [t.contiguous() for t in (gh, bh, gw, bw, gc, bc)]
        y = torch.empty_like(x)
        gstat = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        ws = _workspace(nbytes, x.device)
        args = TnFwdArgs(desc, ptr(x), *[ptr(t) for t in ps], ptr(y), ptr(gstat), ws.data_ptr(), ws.numel())
        check(_call("kmu_triplenorm_fwd", (B, Cc, HW), lib.kmu_triplenorm_fwd, C.byref(args), stream_ptr()), "kmu_triplenorm_fwd")
        ctx.save_for_backward(x, gstat, ps[0], ps[2], ps[4])
        ctx.desc = desc
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, gstat, gh, gw, gc = ctx.saved_tensors
        desc = ctx.desc
        dy = dy.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        grads = [torch.empty_like(gh) for _ in range(6)]
        ws = _workspace(lib.kmu_triplenorm_workspace_bytes(C.byref(desc)), x.device)
        args = TnBwdArgs(desc, ptr(x), ptr(dy), ptr(gstat), ptr(gh), ptr(gw), ptr(gc), ptr(dx), *[ptr(t) for t in grads], ws.data_ptr(),
                         ws.numel())
        check(_call("kmu_triplenorm_bwd", (desc.B, desc.C, desc.HW), lib.kmu_triplenorm_bwd, C.byref(args), stream_ptr()),
              "kmu_triplenorm_bwd")
        return (dx, *grads, None, None)


def triplenorm_supported(channels):
    return channels in (16, 32, 64)


def triplenorm(x, gh, bh, gw, bw, gc, bc, eps_gn=1e-5, eps_ln=1e-5):
    """TripleNorm.forward (KM_UNetV3_SH.py:277-284): (GN1(x; gh,bh) + GN1(x; gw,bw) + LayerNorm over C (gc,bc)) / 3."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.triplenorm: CUDA tensors only (no CPU fallback)")
    return _TripleNormFn.apply(x, gh, bh, gw, bw, gc, bc, eps_gn, eps_ln)


class _QkvGateFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, qkv):
        lib = _lib.lib()
        qkv = qkv.contiguous()
        B, C3 = qkv.shape[0], qkv.shape[1]
        Cc = C3 // 3
        HW = qkv.numel() // (B * C3)
        out = torch.empty((B, Cc) + tuple(qkv.shape[2:]), dtype=torch.float32, device=qkv.device)
        check(_call("kmu_qkv_gate_fwd", (B, Cc, HW), lib.kmu_qkv_gate_fwd, ptr(qkv), ptr(out), B, Cc, HW, stream_ptr()), "kmu_qkv_gate_fwd")
        ctx.save_for_backward(qkv)
        ctx.dims = (B, Cc, HW)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        lib = _lib.lib()
        (qkv,) = ctx.saved_tensors
        B, Cc, HW = ctx.dims
        dout = dout.to(torch.float32).contiguous()
        dqkv = torch.empty_like(qkv)
        check(_call("kmu_qkv_gate_bwd", (B, Cc, HW), lib.kmu_qkv_gate_bwd, ptr(qkv), ptr(dout), ptr(dqkv), B, Cc, HW, stream_ptr()),
              "kmu_qkv_gate_bwd")
        return dqkv


def qkv_gate(qkv):
    """DirectionAttention's gate (KM_UNetV3_SH.py:259-261): sigmoid(q * k) * v on the channel thirds of qkv (B,3C,H,W)."""
    if not qkv.is_cuda:
        raise RuntimeError("km_unet_b200.qkv_gate: CUDA tensors only (no CPU fallback)")
    return _QkvGateFn.apply(qkv)


class _LerpMixFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, m, alpha):
        lib = _lib.lib()
        x, m, al = x.contiguous(), m.contiguous(), alpha.contiguous()
        B, Cc = x.shape[0], x.shape[1]
        HW = x.numel() // (B * Cc)
        y = torch.empty_like(x)
        check(_call("kmu_lerpmix_fwd", (B, Cc, HW), lib.kmu_lerpmix_fwd, ptr(x), ptr(m), ptr(al), ptr(y), B, Cc, HW, stream_ptr()),
              "kmu_lerpmix_fwd")
        ctx.save_for_backward(x, m, al)
        ctx.dims = (B, Cc, HW)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, m, al = ctx.saved_tensors
        B, Cc, HW = ctx.dims
        dy = dy.to(torch.float32).contiguous()
        dx, dm, dal = torch.empty_like(x), torch.empty_like(m), torch.empty_like(al)
        ws = _workspace(lib.kmu_lerpmix_bwd_workspace_bytes(B, Cc, HW), x.device)
        check(_call("kmu_lerpmix_bwd", (B, Cc, HW), lib.kmu_lerpmix_bwd, ptr(x), ptr(m), ptr(dy), ptr(al), ptr(dx), ptr(dm), ptr(dal), B, Cc,
                    HW, ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_lerpmix_bwd")
        return dx, dm, dal


def lerpmix_supported(x):
    return x.is_cuda and x.dim() >= 3 and (x.numel() // (x.shape[0] * x.shape[1])) % 4 == 0


def lerpmix(x, m, alpha):
    """(1 - sigmoid(alpha_c)) x + sigmoid(alpha_c) m with a raw per-channel alpha (C,): EfficientViMBlock's mixer layer-scale."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.lerpmix: CUDA tensors only (no CPU fallback)")
    return _LerpMixFn.apply(x, m.reshape(x.shape), alpha)


class _GroupNormFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, groups, eps):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc = x.shape[0], x.shape[1]
        HW = x.numel() // (B * Cc)
        y = torch.empty_like(x)
        mean = torch.empty(B, groups, dtype=torch.float32, device=x.device)
        rstd = torch.empty(B, groups, dtype=torch.float32, device=x.device)
        ws = _workspace(lib.kmu_groupnorm_fwd_workspace_bytes(B, Cc, HW, groups), x.device)
        w = None if weight is None else weight.contiguous()
        b = None if bias is None else bias.contiguous()
        check(_call("kmu_groupnorm_fwd", (B, Cc, HW, groups), lib.kmu_groupnorm_fwd, ptr(x), ptr(w), ptr(b), ptr(y), ptr(mean), ptr(rstd),
                    B, Cc, HW, groups, float(eps), ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_groupnorm_fwd")
        ctx.save_for_backward(x, w, mean, rstd)
        ctx.dims = (B, Cc, HW, groups)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        B, Cc, HW, groups = ctx.dims
        mask = [ctx.needs_input_grad[0], w is not None and ctx.needs_input_grad[1], w is not None and ctx.needs_input_grad[2]]
        dx, dw, db = torch.ops.aten.native_group_norm_backward(dy.contiguous(), x, mean, rstd, w, B, Cc, HW, groups, mask)
        return dx, dw, db, None, None


def groupnorm_supported(x):
    return x.is_cuda and x.dim() >= 3 and (x.numel() // (x.shape[0] * x.shape[1])) % 4 == 0


def groupnorm(x, weight, bias, groups, eps=1e-5):
    """nn.GroupNorm forward with split statistics CTAs; backward = the library's native_group_norm_backward on the saved mean / rstd."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.groupnorm: CUDA tensors only (no CPU fallback)")
    return _GroupNormFn.apply(x, weight, bias, int(groups), float(eps))


class _ResizeBilinearAcFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, size):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        OH, OW = int(size[0]), int(size[1])
        out = torch.empty(B, Cc, OH, OW, dtype=torch.float32, device=x.device)
        check(_call("kmu_resize_bilinear_ac_fwd", (B, Cc, H, W, OH, OW), lib.kmu_resize_bilinear_ac_fwd, ptr(x), ptr(out), B * Cc, H, W,
                    OH, OW, stream_ptr()), "kmu_resize_bilinear_ac_fwd")
        ctx.in_size, ctx.out_size = (B, Cc, H, W), (OH, OW)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        # the library's own backward kernel (one thread per gradient element) is adequate; only the forward needed replacing
        dx = torch.ops.aten.upsample_bilinear2d_backward(dout.contiguous(), list(ctx.out_size), list(ctx.in_size), True, None, None)
        return dx, None


def resize_bilinear_ac(x, size):
    """F.interpolate(x, size=size, mode='bilinear', align_corners=True) (the skip-connection resizes of KM_UNetV3.forward)."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.resize_bilinear_ac: CUDA tensors only (no CPU fallback)")
    return _ResizeBilinearAcFn.apply(x, tuple(size))


class _Combine3Fn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, f0, f1, f2, coef):
        lib = _lib.lib()
        x, f0, f1, f2 = x.contiguous(), f0.contiguous(), f1.contiguous(), f2.contiguous()
        B = x.shape[0]
        n = x.numel() // B
        coef = coef.reshape(B, 3).contiguous()
        out = torch.empty_like(x)
        check(_call("kmu_combine3_fwd", (B, n), lib.kmu_combine3_fwd, ptr(x), ptr(f0), ptr(f1), ptr(f2), ptr(coef), ptr(out), B, n,
                    stream_ptr()), "kmu_combine3_fwd")
        ctx.save_for_backward(f0, f1, f2, coef)
        ctx.dims = (B, n)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        f0, f1, f2, coef = ctx.saved_tensors
        B, n = ctx.dims
        dy = dy.to(torch.float32).contiguous()
        df0, df1, df2 = torch.empty_like(f0), torch.empty_like(f1), torch.empty_like(f2)
        dcoef = torch.empty_like(coef)
        ws = _workspace(lib.kmu_combine3_bwd_workspace_bytes(B, n), dy.device)
        check(_call("kmu_combine3_bwd", (B, n), lib.kmu_combine3_bwd, ptr(dy), ptr(f0), ptr(f1), ptr(f2), ptr(coef), ptr(df0), ptr(df1),
                    ptr(df2), ptr(dcoef), B, n, ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_combine3_bwd")
        return dy, df0, df1, df2, dcoef


def combine3_supported(x):
    return x.is_cuda and (x.numel() // x.shape[0]) % 4 == 0


def combine3(x, f0, f1, f2, coef):
    """x + sum_i coef[:, i] * f_i with per-sample coefficients coef (B,3): EnhancedViMBlock's gated fusion + DropPath + residual."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.combine3: CUDA tensors only (no CPU fallback)")
    return _Combine3Fn.apply(x, f0, f1, f2, coef)


class _SmallConvFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cin, H, W = x.shape
        Cout, _, kh, kw = weight.shape
        desc = ScDesc(B, Cin, Cout, H, W, kh, kw)
        w = weight.contiguous()
        b = None if bias is None else bias.contiguous()
        y = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x.device)
        check(_call("kmu_smallconv_fwd", (B, Cin, Cout, H, W, kh, kw), lib.kmu_smallconv_fwd, C.byref(desc), ptr(x), ptr(w), ptr(b), ptr(y),
                    stream_ptr()), "kmu_smallconv_fwd")
        ctx.save_for_backward(x, w)
        ctx.desc, ctx.has_bias = desc, bias is not None
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        lib = _lib.lib()
        x, w = ctx.saved_tensors
        d = ctx.desc
        dy = dy.to(torch.float32).contiguous()
        need_x = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty(d.Cout, dtype=torch.float32, device=x.device) if (need_w and ctx.has_bias) else None
        ws = _workspace(lib.kmu_smallconv_bwd_workspace_bytes(C.byref(d)), x.device)
        check(_call("kmu_smallconv_bwd", (d.B, d.Cin, d.Cout, d.H, d.W, d.kh, d.kw), lib.kmu_smallconv_bwd, C.byref(d), ptr(x), ptr(dy),
                    ptr(w), ptr(dx), ptr(dw), ptr(db), ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_smallconv_bwd")
        return dx, dw, db


def smallconv_supported(cin, cout, kh, kw):
    d = ScDesc(1, int(cin), int(cout), 1, 1, int(kh), int(kw))
    return bool(_lib.lib().kmu_smallconv_supported(C.byref(d)))


def smallconv(x, weight, bias=None):
    """Dense stride-1 'same' convolution with <= 9 taps (1x3, 3x1, 3x3): weight (Cout,Cin,kh,kw), zero padding k//2."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.smallconv: CUDA tensors only (no CPU fallback)")
    return _SmallConvFn.apply(x, weight, bias)


class _IwpFn(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias):
        lib = _lib.lib()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        w = weight.reshape(Cc, Cc + 1).contiguous()
        b = bias.contiguous()
        out = torch.empty(B, Cc, H // 2, W // 2, dtype=torch.float32, device=x.device)
        check(_call("kmu_iwp_fwd", (B, Cc, H, W), lib.kmu_iwp_fwd, ptr(x), ptr(w), ptr(b), ptr(out), B, Cc, H, W, stream_ptr()), "kmu_iwp_fwd")
        ctx.save_for_backward(x, w)
        ctx.wshape = weight.shape
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        lib = _lib.lib()
        x, w = ctx.saved_tensors
        B, Cc, H, W = x.shape
        dout = dout.to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        db = torch.empty(Cc, dtype=torch.float32, device=x.device)
        ws = _workspace(lib.kmu_iwp_bwd_workspace_bytes(B, Cc, H, W), x.device)
        check(_call("kmu_iwp_bwd", (B, Cc, H, W), lib.kmu_iwp_bwd, ptr(x), ptr(w), ptr(dout), ptr(dx), ptr(dw), ptr(db), B, Cc, H, W,
                    ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_iwp_bwd")
        return dx, dw.reshape(ctx.wshape), db


def iwp_supported(channels, H, W):
    return channels in (16, 32, 64) and H % 2 == 0 and W % 2 == 0


def iwp(x, fusion_weight, fusion_bias):
    """IntelligentWaveletPoolingModule.forward (WPL/iwp.py:116-132): Haar 2x2 analysis + 1x1 fusion, (B,C,H,W) -> (B,C,H/2,W/2)."""
    if not x.is_cuda:
        raise RuntimeError("km_unet_b200.iwp: CUDA tensors only (no CPU fallback)")
    return _IwpFn.apply(x, fusion_weight, fusion_bias)
