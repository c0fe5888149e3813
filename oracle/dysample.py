"""Oracle restatement of DySample (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows the reference:
    _init_pos        DySample_md.py:45-47   (init_pos)
    sample           DySample_md.py:49-61   (sample: meshgrid + normalise + pixel_shuffle + grid_sample collapse to
                                             "source index = w + off_x, h + off_y, clamped, bilinear")
    forward_lp       DySample_md.py:63-68   (dysample_lp)
The index form was verified against the reference's grid_sample pipeline (tests/test_oracle_dysample.py).
Pure torch-CPU arithmetic, fp32 or fp64.
"""
import torch
import torch.nn.functional as F


def init_pos(scale=2, groups=4, dtype=torch.float32):
    """(1, 2*groups*scale^2, 1, 1); channel d*G*s^2 + g*s^2 + i*s + j holds the x (d=0, depends on j) or
    y (d=1, depends on i) sub-pixel start."""
    h = (torch.arange(scale, dtype=dtype) - (scale - 1) / 2) / scale
    pos = torch.empty(2, groups, scale, scale, dtype=dtype)
    pos[0] = h.reshape(1, 1, scale)
    pos[1] = h.reshape(1, scale, 1)
    return pos.reshape(1, -1, 1, 1)


def sample(x, offset, scale=2, groups=4):
    """x (B,C,H,W), offset (B, 2*groups*scale^2, H, W) in input-pixel units -> (B,C,scale*H,scale*W)."""
    B, C, H, W = x.shape
    s, G = scale, groups
    Cg = C // G
    off = offset.reshape(B, 2, G, s, s, H, W)
    ww = torch.arange(W, dtype=x.dtype).reshape(1, 1, 1, 1, 1, W)
    hh = torch.arange(H, dtype=x.dtype).reshape(1, 1, 1, 1, H, 1)
    sx = (ww + off[:, 0]).clamp(0, W - 1)                     # (B,G,s,s,H,W)
    sy = (hh + off[:, 1]).clamp(0, H - 1)
    x0 = sx.floor()
    y0 = sy.floor()
    fx = sx - x0
    fy = sy - y0
    x0 = x0.long()
    y0 = y0.long()
    xg = x.reshape(B, G, Cg, H * W)

    def tap(yy, xx):
        ok = ((yy < H) & (xx < W)).to(x.dtype)               # the +1 neighbour past the edge contributes 0
        idx = (yy.clamp(max=H - 1) * W + xx.clamp(max=W - 1)).reshape(B, G, 1, -1).expand(B, G, Cg, -1)
        v = torch.gather(xg, 3, idx).reshape(B, G, Cg, s, s, H, W)
        return v * ok.unsqueeze(2)

    fxe, fye = fx.unsqueeze(2), fy.unsqueeze(2)
    out = (tap(y0, x0) * (1 - fxe) * (1 - fye) + tap(y0, x0 + 1) * fxe * (1 - fye)
           + tap(y0 + 1, x0) * (1 - fxe) * fye + tap(y0 + 1, x0 + 1) * fxe * fye)        # (B,G,Cg,s,s,H,W)
    return out.permute(0, 1, 2, 5, 3, 6, 4).reshape(B, C, s * H, s * W)


def dysample_lp(x, offset_weight, offset_bias, init_pos_buf, scale=2, groups=4):
    off = F.conv2d(x, offset_weight, offset_bias) * 0.25 + init_pos_buf
    return sample(x, off, scale, groups)


def sample_grads(x, offset, dout, scale=2, groups=4):
    """Closed-form backward of `sample`: returns (dx, doffset).  The offset gradient is zeroed where the
    coordinate was clipped (including exactly on the border), as torch's grid_sample does."""
    B, C, H, W = x.shape
    s, G = scale, groups
    Cg = C // G
    off = offset.reshape(B, 2, G, s, s, H, W)
    ww = torch.arange(W, dtype=x.dtype).reshape(1, 1, 1, 1, 1, W)
    hh = torch.arange(H, dtype=x.dtype).reshape(1, 1, 1, 1, H, 1)
    rx = ww + off[:, 0]
    ry = hh + off[:, 1]
    sx, sy = rx.clamp(0, W - 1), ry.clamp(0, H - 1)
    inx = ((rx > 0) & (rx < W - 1)).to(x.dtype)
    iny = ((ry > 0) & (ry < H - 1)).to(x.dtype)
    x0, y0 = sx.floor(), sy.floor()
    fx, fy = (sx - x0).unsqueeze(2), (sy - y0).unsqueeze(2)
    x0, y0 = x0.long(), y0.long()
    xg = x.reshape(B, G, Cg, H * W)
    dog = dout.reshape(B, G, Cg, H, s, W, s).permute(0, 1, 2, 4, 6, 3, 5)               # (B,G,Cg,s,s,H,W)
    dx = torch.zeros_like(xg)

    def tap(yy, xx, wgt):
        ok = ((yy < H) & (xx < W)).to(x.dtype).unsqueeze(2)
        idx = (yy.clamp(max=H - 1) * W + xx.clamp(max=W - 1)).reshape(B, G, 1, -1).expand(B, G, Cg, -1)
        v = torch.gather(xg, 3, idx).reshape(B, G, Cg, s, s, H, W) * ok
        dx.scatter_add_(3, idx, (dog * wgt * ok).reshape(B, G, Cg, -1))
        return v

    v00 = tap(y0, x0, (1 - fx) * (1 - fy))
    v01 = tap(y0, x0 + 1, fx * (1 - fy))
    v10 = tap(y0 + 1, x0, (1 - fx) * fy)
    v11 = tap(y0 + 1, x0 + 1, fx * fy)
    gx = (((v01 - v00) * (1 - fy) + (v11 - v10) * fy) * dog).sum(2) * inx
    gy = (((v10 - v00) * (1 - fx) + (v11 - v01) * fx) * dog).sum(2) * iny
    doff = torch.stack([gx, gy], dim=1).reshape(B, 2 * G * s * s, H, W)
    return dx.reshape(B, C, H, W), doff
