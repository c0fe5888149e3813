"""oracle/ -- CPU restatement of the KM-UNet hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain torch-CPU / numpy arithmetic, what the reference's
PyTorch modules compute on the hot path:

    kan.py       KANLinear / KANConv2d          convKAN/KANlayers.py:505-660, convKAN/KANConv2Dlayers.py:5-37
    hsmssd.py    HSMSSD, LayerNorm1D, EfficientViMBlock
                                                 vim_block_init/efficient_vim_init.py:14-97, vim_utils_init.py:34-130
    dysample.py  DySample ('lp', scale 2)        DySample_md.py:20-81
    dagem.py     DAGEM                           DAGEM_md.py:7-111

Parity pin: the restatements are checked (tests/test_oracle_*.py) against
  (1) golden vectors under tests/golden/*.npz, produced by importing the UNMODIFIED
      reference modules from /root/reference (tests/golden/make_golden.py), and
  (2) the live reference modules whenever /root/reference is present (this container).
The reference ships no tests / known-answer vectors of its own, so (1)/(2)
are the pin.  Third-party arithmetic that is not in /root/reference (torchvision
DeformConv2d, torchmetrics SSIM) is "parity unpinned": see DESIGN.md.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this package.  The product (km_unet_b200/) never does: it raises if the CUDA
extension is missing instead of falling back to anything here.
"""
