"""Oracle restatement of KANLinear / KANConv2d (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows the reference:
    knots                convKAN/KANlayers.py:526-535   (make_grid)
    Cox-de Boor basis    convKAN/KANlayers.py:577-610   (bspline_basis)
    layer forward        convKAN/KANlayers.py:644-660   (kan_linear)
    im2col wrapper       convKAN/KANConv2Dlayers.py:15-37 (kanconv2d)
Closed-form gradients (kan_linear_grads, kanconv2d_grads) are what the CUDA backward implements;
they are checked against autograd of the reference module in tests/test_oracle_kan.py.

Pure torch-CPU arithmetic; works in fp32 or fp64 (dtype follows the inputs).
"""
import torch
import torch.nn.functional as F


def make_grid(in_features, grid_size=5, spline_order=3, grid_range=(-1.0, 1.0), dtype=torch.float32):
    """(in_features, grid_size + 2*spline_order + 1) knot table; fp32-rounded like the reference buffer."""
    h = (grid_range[1] - grid_range[0]) / grid_size
    knots = torch.arange(-spline_order, grid_size + spline_order + 1) * h + grid_range[0]
    return knots.to(dtype).expand(in_features, -1).contiguous()


def bspline_basis(x, grid, spline_order=3):
    """x (M, in), grid (in, T) -> (M, in, T - spline_order - 1), half-open order-0 indicators then
    `spline_order` Cox-de Boor levels (vectorised over the knot axis)."""
    assert x.dim() == 2 and x.shape[1] == grid.shape[0]
    g = grid.to(x.dtype)
    xe = x[:, :, None]
    level = ((xe >= g[:, :-1]) & (xe < g[:, 1:])).to(x.dtype)
    for p in range(1, spline_order + 1):
        span_lo = g[:, p:-1] - g[:, :-(p + 1)]
        span_hi = g[:, p + 1:] - g[:, 1:-p]
        level = (xe - g[:, :-(p + 1)]) / span_lo * level[..., :-1] + (g[:, p + 1:] - xe) / span_hi * level[..., 1:]
    return level


def bspline_basis_and_derivative(x, grid, spline_order=3):
    """Basis values and d/dx of each basis function (recursive derivative formula
    B'_{j,p} = p/(t_{j+p}-t_j) B_{j,p-1} - p/(t_{j+p+1}-t_{j+1}) B_{j+1,p-1})."""
    g = grid.to(x.dtype)
    p = spline_order
    lower = bspline_basis_level(x, g, p - 1)          # (M, in, T-p)
    full = bspline_basis(x, g, p)
    nb = full.shape[-1]
    cols = []
    for j in range(nb):
        a = p / (g[:, j + p] - g[:, j])
        b = p / (g[:, j + p + 1] - g[:, j + 1])
        cols.append(a * lower[..., j] - b * lower[..., j + 1])
    return full, torch.stack(cols, dim=-1)


def bspline_basis_level(x, grid, order):
    if order == 0:
        T = grid.shape[1]
        return torch.stack([((x >= grid[:, j]) & (x < grid[:, j + 1])).to(x.dtype) for j in range(T - 1)], dim=-1)
    return bspline_basis(x, grid, order)


def uniform_cubic_basis(x, t0, h, n_basis=8):
    """Closed form used by the tensor-core producer: uniform knots t_j = t0 + j*h, cubic.
    Returns (M, in, n_basis).  Equals bspline_basis up to the fp32 rounding of the knot table."""
    s = (x - t0) / h
    i = torch.floor(s)
    u = s - i
    w = torch.stack([(1 - u) ** 3, 3 * u ** 3 - 6 * u ** 2 + 4, -3 * u ** 3 + 3 * u ** 2 + 3 * u + 1, u ** 3], dim=-1) / 6
    out = x.new_zeros(x.shape + (n_basis,))
    ii = i.long()
    for r in range(4):
        j = ii - 3 + r
        ok = (j >= 0) & (j < n_basis) & (s >= 0) & (s < n_basis + 3)
        out.scatter_add_(-1, j.clamp(0, n_basis - 1).unsqueeze(-1), (w[..., r] * ok.to(x.dtype)).unsqueeze(-1))
    return out


def kan_linear(x, base_weight, spline_weight, spline_scaler, grid, spline_order=3):
    """y = SiLU(x) @ Wb^T + vec(B(x)) @ (Ws * s[...,None]).view(out,-1)^T ; x (M,in) -> (M,out)."""
    out_f = base_weight.shape[0]
    base = F.silu(x) @ base_weight.t()
    w_eff = spline_weight if spline_scaler is None else spline_weight * spline_scaler[..., None]
    basis = bspline_basis(x, grid, spline_order)
    return base + basis.reshape(x.shape[0], -1) @ w_eff.reshape(out_f, -1).t()


def unfold_patches(x, kernel_size, stride, padding):
    """(B,C,H,W) -> (B*Ho*Wo, C*k*k) with feature index c*k*k + ki*k + kj and zero padding (im2col)."""
    B, C, H, W = x.shape
    k = kernel_size
    Ho = (H + 2 * padding - k) // stride + 1
    Wo = (W + 2 * padding - k) // stride + 1
    cols = F.unfold(x, kernel_size=k, stride=stride, padding=padding)          # (B, C*k*k, Ho*Wo)
    return cols.transpose(1, 2).reshape(B * Ho * Wo, C * k * k), (Ho, Wo)


def kanconv2d(x, base_weight, spline_weight, spline_scaler, grid, kernel_size=3, stride=1, padding=0, spline_order=3):
    B = x.shape[0]
    patches, (Ho, Wo) = unfold_patches(x, kernel_size, stride, padding)
    y = kan_linear(patches, base_weight, spline_weight, spline_scaler, grid, spline_order)
    # NB the reference reshapes (B, L, Cout) -> transpose -> (B, Cout, Ho, Wo)
    return y.reshape(B, Ho * Wo, -1).transpose(1, 2).reshape(B, -1, Ho, Wo)


def silu_grad(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


def kan_linear_grads(x, dy, base_weight, spline_weight, spline_scaler, grid, spline_order=3):
    """Closed-form gradients of kan_linear: returns dx, d_base_weight, d_spline_weight, d_spline_scaler."""
    basis, dbasis = bspline_basis_and_derivative(x, grid, spline_order)          # (M,in,nb)
    w_eff = spline_weight * spline_scaler[..., None]                             # (out,in,nb)
    d_base = dy.t() @ F.silu(x)
    d_weff = torch.einsum("mo,mfj->ofj", dy, basis)
    d_spline = d_weff * spline_scaler[..., None]
    d_scaler = (d_weff * spline_weight).sum(-1)
    dx = (dy @ base_weight) * silu_grad(x) + torch.einsum("mo,ofj,mfj->mf", dy, w_eff, dbasis)
    return dx, d_base, d_spline, d_scaler


def fold_patches(dcols, x_shape, kernel_size, stride, padding):
    """Adjoint of unfold_patches (col2im)."""
    B, C, H, W = x_shape
    k = kernel_size
    Ho = (H + 2 * padding - k) // stride + 1
    Wo = (W + 2 * padding - k) // stride + 1
    d = dcols.reshape(B, Ho, Wo, C, k, k)
    dxp = dcols.new_zeros(B, C, H + 2 * padding, W + 2 * padding)
    for ki in range(k):
        for kj in range(k):
            dxp[:, :, ki:ki + stride * Ho:stride, kj:kj + stride * Wo:stride] += d[..., ki, kj].permute(0, 3, 1, 2)
    return dxp[:, :, padding:padding + H, padding:padding + W]


def kanconv2d_grads(x, dy, base_weight, spline_weight, spline_scaler, grid, kernel_size=3, stride=1, padding=0,
                    spline_order=3):
    B, Cout = dy.shape[0], dy.shape[1]
    patches, _ = unfold_patches(x, kernel_size, stride, padding)
    dy2 = dy.reshape(B, Cout, -1).transpose(1, 2).reshape(-1, Cout)
    dcols, d_base, d_spline, d_scaler = kan_linear_grads(patches, dy2, base_weight, spline_weight, spline_scaler, grid,
                                                         spline_order)
    return fold_patches(dcols, x.shape, kernel_size, stride, padding), d_base, d_spline, d_scaler


def phi_expand(x, grid_row, spline_order=3):
    """Per-pixel expansion Phi(x) = [SiLU(x), B_0(x) .. B_{nb-1}(x)] for a grid shared by all features:
    x (B,C,H,W) -> (B,C,1+nb,H,W).  The tensor-core kernels convolve this image."""
    B, C, H, W = x.shape
    flat = x.reshape(-1, 1)
    basis = bspline_basis(flat, grid_row.reshape(1, -1), spline_order)[:, 0]          # (N, nb)
    phi = torch.cat([F.silu(flat), basis], dim=1)
    return phi.reshape(B, C, H, W, -1).permute(0, 1, 4, 2, 3).contiguous()


def kanconv2d_as_phi_conv(x, base_weight, spline_weight, spline_scaler, grid, kernel_size=3, stride=1, padding=0,
                          spline_order=3):
    """KANConv2d == conv2d over the Phi-expanded image padded with Phi(0) (valid when every feature shares one
    knot row).  Restates the implicit-GEMM formulation of the tcgen05 kernel for cross-checking."""
    B, C, H, W = x.shape
    k = kernel_size
    nb = spline_weight.shape[-1]
    Cout = base_weight.shape[0]
    xp = F.pad(x, (padding,) * 4)                       # zero pad first: Phi(0) appears at the border
    phi = phi_expand(xp, grid[0], spline_order)         # (B,C,1+nb,Hp,Wp)
    w_eff = spline_weight * spline_scaler[..., None]    # (out, C*k*k, nb)
    wb = base_weight.reshape(Cout, C, 1, k, k)
    ws = w_eff.reshape(Cout, C, k, k, nb).permute(0, 1, 4, 2, 3)
    w_full = torch.cat([wb, ws], dim=2).reshape(Cout, C * (1 + nb), k, k)
    return F.conv2d(phi.reshape(B, C * (1 + nb), H + 2 * padding, W + 2 * padding), w_full, stride=stride)
