"""KANLinear / KANConv2d with the reference's constructor signature, parameter names and state_dict layout
(convKAN/KANlayers.py:505-731, convKAN/KANConv2Dlayers.py:5-37); forward/backward run in libkmunet.so.

Construction-time helpers (knot table, least-squares spline init, update_grid, regularisation) are host logic and
stay in torch; they are never on the training hot path.
"""
import math

import torch
import torch.nn as nn

from .. import config, ops


class KANLinear(nn.Module):
    def __init__(self, in_features, out_features, grid_size=5, spline_order=3, scale_noise=0.1, scale_base=1.0,
                 scale_spline=1.0, enable_standalone_scale_spline=True, base_activation=torch.nn.SiLU, grid_eps=0.02,
                 grid_range=[-1, 1]):
        super().__init__()
        if base_activation is not torch.nn.SiLU:
            raise NotImplementedError("km_unet_b200.KANLinear implements the SiLU base branch only")
        self.in_features, self.out_features = in_features, out_features
        self.grid_size, self.spline_order = grid_size, spline_order
        step = (grid_range[1] - grid_range[0]) / grid_size
        knots = torch.arange(-spline_order, grid_size + spline_order + 1) * step + grid_range[0]
        self.register_buffer("grid", knots.expand(in_features, -1).contiguous())
        n_basis = grid_size + spline_order
        self.base_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.spline_weight = nn.Parameter(torch.empty(out_features, in_features, n_basis))
        if enable_standalone_scale_spline:
            self.spline_scaler = nn.Parameter(torch.empty(out_features, in_features))
        self.scale_noise, self.scale_base, self.scale_spline = scale_noise, scale_base, scale_spline
        self.enable_standalone_scale_spline = enable_standalone_scale_spline
        self.base_activation = base_activation()
        self.grid_eps = grid_eps
        self.precision = None          # None -> km_unet_b200.config.kan_precision
        self.reset_parameters()
        # load_state_dict rewrites the knot table: refresh the cached knot summary right away (a device -> host read) so that
        # the next forward -- possibly under CUDA-graph capture, where such a read is illegal -- finds it valid
        self.register_load_state_dict_post_hook(KANLinear._refresh_grid_meta)

    # -- host-side helpers (init / grid adaptation only) -------------------------------------------------------
    def reset_parameters(self):
        # same RNG consumption order as the reference (KANlayers.py:555-575), so a seed reproduces its init
        nn.init.kaiming_uniform_(self.base_weight, a=math.sqrt(5) * self.scale_base)
        with torch.no_grad():
            noise = (torch.rand(self.grid_size + 1, self.in_features, self.out_features) - 0.5) \
                * self.scale_noise / self.grid_size
            inner_knots = self.grid.T[self.spline_order:-self.spline_order]
            coeff = self.curve2coeff(inner_knots, noise)
            self.spline_weight.data.copy_(coeff if self.enable_standalone_scale_spline else self.scale_spline * coeff)
            if self.enable_standalone_scale_spline:
                nn.init.kaiming_uniform_(self.spline_scaler, a=math.sqrt(5) * self.scale_spline)

    def b_splines(self, x):
        """(M, in) -> (M, in, grid_size + spline_order) Cox-de Boor values (host helper; the hot path evaluates the
        same recursion inside the CUDA kernels)."""
        assert x.dim() == 2 and x.size(1) == self.in_features
        t = self.grid
        xe = x.unsqueeze(-1)
        b = ((xe >= t[:, :-1]) & (xe < t[:, 1:])).to(x.dtype)
        for p in range(1, self.spline_order + 1):
            lo = (xe - t[:, :-(p + 1)]) / (t[:, p:-1] - t[:, :-(p + 1)])
            hi = (t[:, p + 1:] - xe) / (t[:, p + 1:] - t[:, 1:-p])
            b = lo * b[:, :, :-1] + hi * b[:, :, 1:]
        assert b.shape == (x.size(0), self.in_features, self.grid_size + self.spline_order)
        return b.contiguous()

    def curve2coeff(self, x, y):
        """Least-squares spline coefficients interpolating y (M, in, out) at x (M, in) -> (out, in, n_basis)."""
        assert x.dim() == 2 and x.size(1) == self.in_features
        assert y.shape == (x.size(0), self.in_features, self.out_features)
        design = self.b_splines(x).transpose(0, 1)
        sol = torch.linalg.lstsq(design, y.transpose(0, 1)).solution
        out = sol.permute(2, 0, 1)
        assert out.shape == (self.out_features, self.in_features, self.grid_size + self.spline_order)
        return out.contiguous()

    @property
    def scaled_spline_weight(self):
        if self.enable_standalone_scale_spline:
            return self.spline_weight * self.spline_scaler.unsqueeze(-1)
        return self.spline_weight

    def _precision(self):
        return config.precision_code(self.precision)

    @staticmethod
    def _refresh_grid_meta(module, incompatible_keys):
        if module.grid.is_cuda:
            module._grid_meta()

    def _grid_meta(self):
        """Cached (uniform_and_shared, t0, h) of the knot table, refreshed when the buffer is modified or moved."""
        key = (self.grid._version, self.grid.data_ptr())
        if getattr(self, "_grid_meta_key", None) != key:
            self._grid_meta_val = ops.grid_info(self.grid)
            self._grid_meta_key = key
        return self._grid_meta_val

    # -- hot path ----------------------------------------------------------------------------------------------
    def forward(self, x):
        assert x.dim() == 2 and x.size(1) == self.in_features
        scaler = self.spline_scaler if self.enable_standalone_scale_spline else None
        return ops.kanlinear(x, self.base_weight, self.spline_weight, scaler, self.grid, self.grid_size, self.spline_order,
                             self._precision(), self._grid_meta())

    @torch.no_grad()
    def update_grid(self, x, margin=0.01):
        assert x.dim() == 2 and x.size(1) == self.in_features
        n = x.size(0)
        basis = self.b_splines(x).permute(1, 0, 2)
        coeff = self.scaled_spline_weight.permute(1, 2, 0)
        y = torch.bmm(basis, coeff).permute(1, 0, 2)
        xs = torch.sort(x, dim=0)[0]
        pick = torch.linspace(0, n - 1, self.grid_size + 1, dtype=torch.int64, device=x.device)
        adaptive = xs[pick]
        step = (xs[-1] - xs[0] + 2 * margin) / self.grid_size
        uniform = torch.arange(self.grid_size + 1, dtype=torch.float32, device=x.device).unsqueeze(1) * step + xs[0] - margin
        g = self.grid_eps * uniform + (1 - self.grid_eps) * adaptive
        k = self.spline_order
        left = g[:1] - step * torch.arange(k, 0, -1, device=x.device).unsqueeze(1)
        right = g[-1:] + step * torch.arange(1, k + 1, device=x.device).unsqueeze(1)
        self.grid.copy_(torch.cat([left, g, right], dim=0).T)
        self.spline_weight.data.copy_(self.curve2coeff(x, y))

    def regularization_loss(self, regularize_activation=1.0, regularize_entropy=1.0):
        l1 = self.spline_weight.abs().mean(-1)
        total = l1.sum()
        p = l1 / total
        entropy = -torch.sum(p * p.log())
        return regularize_activation * total + regularize_entropy * entropy


class KANConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.kanlayer = KANLinear(in_channels * kernel_size * kernel_size, out_channels)

    def forward(self, x):
        assert x.size(1) == self.in_channels
        k = self.kanlayer
        scaler = k.spline_scaler if k.enable_standalone_scale_spline else None
        return ops.kanconv2d(x, k.base_weight, k.spline_weight, scaler, k.grid, self.kernel_size, self.stride, self.padding,
                             k.grid_size, k.spline_order, k._precision(), k._grid_meta())


KAN_Convolutional_Layer = KANConv2d  # name used by BASELINE.json's north_star
