"""Oracle restatement of DAGEM (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows the reference DAGEM_md.py:56-111 (forward) with the parameter names of DAGEM_md.py:15-54.
`deform_conv2d` restates torchvision.ops.deform_conv2d (torchvision is a third-party dependency of the
reference -- pinned 0.14.0 in requirements.txt:83, call sites DAGEM_md.py:4,46,101 -- whose source is not in
/root/reference: parity for that op is pinned on the torchvision CPU operator installed here, see
tests/test_oracle_dagem.py, and is "unpinned" with respect to the reference's own tree).
Pure torch-CPU arithmetic, fp32 or fp64.
"""
import torch
import torch.nn.functional as F


def _bn_rows(v, weight, bias, running_mean, running_var, training, eps=1e-5, momentum=0.1, update=False):
    """BatchNorm1d over the rows of v (R, F).  update=True also moves the running statistics as nn.BatchNorm1d does in
    train mode (momentum 0.1, UNBIASED batch variance; DAGEM_md.py:17,24,31,38,52 build them with the defaults)."""
    if training:
        mu = v.mean(dim=0)
        var = ((v - mu) ** 2).mean(dim=0)
        if update and running_mean is not None:
            n = v.shape[0]
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(momentum * mu.detach().to(running_mean.dtype))
                running_var.mul_(1 - momentum).add_(momentum * (var.detach() * n / max(n - 1, 1)).to(running_var.dtype))
    else:
        mu, var = running_mean, running_var
    return (v - mu) / torch.sqrt(var + eps) * weight + bias


def deform_conv2d(x, offset, weight, bias=None, padding=1):
    """3x3-style deformable convolution, stride 1, dilation 1, one offset group.
    offset (B, 2*kh*kw, Ho, Wo) ordered (dy, dx) per tap; bilinear sampling with zeros outside the image."""
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    Ho, Wo = offset.shape[2], offset.shape[3]
    oy = torch.arange(Ho, dtype=x.dtype).reshape(1, Ho, 1)
    ox = torch.arange(Wo, dtype=x.dtype).reshape(1, 1, Wo)
    flat = x.reshape(B, C, H * W)
    cols = []
    for t in range(kh * kw):
        ki, kj = divmod(t, kw)
        py = oy - padding + ki + offset[:, 2 * t]
        px = ox - padding + kj + offset[:, 2 * t + 1]
        inside = ((py > -1) & (py < H) & (px > -1) & (px < W)).to(x.dtype)
        y0, x0 = py.floor(), px.floor()
        ly, lx = py - y0, px - x0
        y0, x0 = y0.long(), x0.long()
        acc = x.new_zeros(B, C, Ho, Wo)
        for dy, dx, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
            yy, xx = y0 + dy, x0 + dx
            ok = ((yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)).to(x.dtype) * inside
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).reshape(B, 1, -1).expand(B, C, -1)
            acc = acc + torch.gather(flat, 2, idx).reshape(B, C, Ho, Wo) * (wgt * ok).unsqueeze(1)
        cols.append(acc)
    col = torch.stack(cols, dim=2)                                   # (B, C, kh*kw, Ho, Wo)
    out = torch.einsum("oct,bcthw->bohw", weight.reshape(Co, C, kh * kw), col)
    if bias is not None:
        out = out + bias.reshape(1, -1, 1, 1)
    return out


def dagem(x, P, training=True):
    """x (B,C,H,W); P = state_dict of the module (keys like 'edge_aggregation_func.0.weight')."""
    B, C, H, W = x.shape
    Ch = C // 2

    def bn(prefix, v):
        return _bn_rows(v, P[prefix + ".weight"], P[prefix + ".bias"], P[prefix + ".running_mean"],
                        P[prefix + ".running_var"], training)

    nbrs = torch.stack([torch.roll(x, 1, 2), torch.roll(x, -1, 2), torch.roll(x, 1, 3), torch.roll(x, -1, 3)], dim=-1)
    edge = nbrs * x.unsqueeze(-1)                                                    # (B,C,H,W,4)
    agg = edge.reshape(-1, 4) @ P["edge_aggregation_func.0.weight"].t() + P["edge_aggregation_func.0.bias"]
    agg = F.relu(bn("edge_aggregation_func.1", agg)).reshape(B, C, H, W)
    vfeat = torch.cat([x, agg], dim=1).permute(0, 2, 3, 1).reshape(-1, 2 * C)
    uv = vfeat @ P["vertex_update_func.0.weight"].t() + P["vertex_update_func.0.bias"]
    uv = F.relu(bn("vertex_update_func.1", uv)).reshape(B, H, W, Ch).permute(0, 3, 1, 2)
    efeat = torch.cat([x.unsqueeze(-1).expand(B, C, H, W, 4), edge], dim=1).permute(0, 2, 3, 4, 1).reshape(-1, 2 * C)
    ue = efeat @ P["edge_update_func.0.weight"].t() + P["edge_update_func.0.bias"]
    ue = F.relu(bn("edge_update_func.1", ue)).reshape(B, H, W, 4, Ch).permute(0, 4, 1, 2, 3).reshape(-1, 4)
    ur = ue @ P["update_edge_reduce_func.0.weight"].t() + P["update_edge_reduce_func.0.bias"]
    ur = F.relu(bn("update_edge_reduce_func.1", ur)).reshape(B, Ch, H, W)
    feat = uv * ur
    offset = F.conv2d(x, P["offset_conv.weight"], P["offset_conv.bias"], padding=1)
    deformed = deform_conv2d(x, offset, P["deform_conv.weight"], P["deform_conv.bias"], padding=1) + x
    z = F.conv2d(torch.cat([deformed, feat], dim=1), P["final_aggregation_layer.0.weight"])
    zr = z.permute(0, 2, 3, 1).reshape(-1, C)
    zr = F.relu(bn("final_aggregation_layer.1", zr))
    return zr.reshape(B, H, W, C).permute(0, 3, 1, 2)


def dagem_gate(x, deformed, lin, bns, training=True):
    """The gating + final aggregation alone (DAGEM_md.py:62-92,104-110) with explicit tensors, mirroring the C ABI
    kmu_dagem_fwd: lin = (ea_w, ea_b, vu_w, vu_b, eu_w, eu_b, er_w, er_b, wf); bns = 5 x (weight, bias, running_mean,
    running_var) in the order edge_aggregation, edge_update, vertex_update, update_edge_reduce, final.  In training the
    running statistics are updated in place, as the reference's BatchNorm modules (and kmu_dagem_fwd) do."""
    ea_w, ea_b, vu_w, vu_b, eu_w, eu_b, er_w, er_b, wf = lin
    B, C, H, W = x.shape
    Ch = C // 2

    def bn(i, v):
        w, b, rm, rv = bns[i]
        return _bn_rows(v, w, b, rm, rv, training, update=True)

    nbrs = torch.stack([torch.roll(x, 1, 2), torch.roll(x, -1, 2), torch.roll(x, 1, 3), torch.roll(x, -1, 3)], dim=-1)
    edge = nbrs * x.unsqueeze(-1)
    agg = F.relu(bn(0, edge.reshape(-1, 4) @ ea_w.t() + ea_b)).reshape(B, C, H, W)
    vfeat = torch.cat([x, agg], dim=1).permute(0, 2, 3, 1).reshape(-1, 2 * C)
    uv = F.relu(bn(2, vfeat @ vu_w.t() + vu_b)).reshape(B, H, W, Ch).permute(0, 3, 1, 2)
    efeat = torch.cat([x.unsqueeze(-1).expand(B, C, H, W, 4), edge], dim=1).permute(0, 2, 3, 4, 1).reshape(-1, 2 * C)
    ue = F.relu(bn(1, efeat @ eu_w.t() + eu_b)).reshape(B, H, W, 4, Ch).permute(0, 4, 1, 2, 3).reshape(-1, 4)
    ur = F.relu(bn(3, ue @ er_w.t() + er_b)).reshape(B, Ch, H, W)
    z = F.conv2d(torch.cat([deformed, uv * ur], dim=1), wf.reshape(C, C + Ch, 1, 1))
    zr = F.relu(bn(4, z.permute(0, 2, 3, 1).reshape(-1, C)))
    return zr.reshape(B, H, W, C).permute(0, 3, 1, 2)
