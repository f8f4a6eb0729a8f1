"""Synthetic reference-format EDVR weights: what bench.py, the tools and the smoke test feed the executor
when no checkpoint is available (there is no network for the released .pth files).

Initialisers restated from the reference: ResidualBlockNoBN kaiming-normal x 0.1 with zero bias
(basicsr/models/archs/arch_util.py:20-48,89-90), DCN weight U(+-1/sqrt(Cin*9)) with zero bias
(basicsr/models/ops/dcn/deform_conv.py:330-337), PyTorch's nn.Conv2d default elsewhere; conv_offset, which the reference
zero-initialises (deform_conv.py:377-381), is drawn from N(0, offset_std^2) so that the sampling is irregular.
oracle/edvr_ref.py carries the same generator for the checker; tests/test_host.py pins the two to identical tensors.
"""
import torch


def make_state_dict(num_feat=64, num_frame=5, deformable_groups=8, num_extract_block=5,
                    num_reconstruct_block=10, with_predeblur=False, hr_in=False,
                    with_tsa=True, seed=0, offset_std=0.02):
    """Synthetic reference-format weights: reference initialisers, except
    conv_offset ~ N(0, offset_std^2) so that the gather is irregular.

    Init rules restated from arch_util.py:20-48,89-90 (ResidualBlockNoBN: kaiming-normal
    x0.1, bias 0), deform_conv.py:330-337 (DCN weight U(+-1/sqrt(Cin*9)), bias 0) and
    PyTorch's nn.Conv2d default for everything else.
    """
    import math
    g = torch.Generator().manual_seed(seed)
    sd = {}
    nf = num_feat

    def conv(key, cout, cin, k, mode="default"):
        fan_in = cin * k * k
        if mode == "res":
            wgt = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan_in) * 0.1
            bias = torch.zeros(cout)
        elif mode == "offset":
            wgt = torch.randn(cout, cin, k, k, generator=g) * offset_std
            bias = torch.randn(cout, generator=g) * offset_std
        else:
            bound = 1.0 / math.sqrt(fan_in)
            wgt = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
            bias = (torch.rand(cout, generator=g) * 2 - 1) * bound
        sd[key + ".weight"], sd[key + ".bias"] = wgt, bias

    def resblock(key):
        conv(key + ".conv1", nf, nf, 3, "res")
        conv(key + ".conv2", nf, nf, 3, "res")

    def dcnpack(key):
        bound = 1.0 / math.sqrt(nf * 9)
        sd[key + ".weight"] = (torch.rand(nf, nf, 3, 3, generator=g) * 2 - 1) * bound
        sd[key + ".bias"] = torch.zeros(nf)
        conv(key + ".conv_offset", deformable_groups * 27, nf, 3, "offset")

    if with_predeblur:
        p = "predeblur."
        conv(p + "conv_first", nf, 3, 3)
        if hr_in:
            conv(p + "stride_conv_hr1", nf, nf, 3)
            conv(p + "stride_conv_hr2", nf, nf, 3)
        conv(p + "stride_conv_l2", nf, nf, 3)
        conv(p + "stride_conv_l3", nf, nf, 3)
        for k in ("resblock_l3", "resblock_l2_1", "resblock_l2_2"):
            resblock(p + k)
        for i in range(5):
            resblock(p + f"resblock_l1.{i}")
        conv("conv_1x1", nf, nf, 1)
    else:
        conv("conv_first", nf, 3, 3)
    for i in range(num_extract_block):
        resblock(f"feature_extraction.{i}")
    for k in ("conv_l2_1", "conv_l2_2", "conv_l3_1", "conv_l3_2"):
        conv(k, nf, nf, 3)
    p = "pcd_align."
    for lvl in (3, 2, 1):
        L = f"l{lvl}"
        conv(p + "offset_conv1." + L, nf, 2 * nf, 3)
        conv(p + "offset_conv2." + L, nf, nf if lvl == 3 else 2 * nf, 3)
        if lvl < 3:
            conv(p + "offset_conv3." + L, nf, nf, 3)
        dcnpack(p + "dcn_pack." + L)
        if lvl < 3:
            conv(p + "feat_conv." + L, nf, 2 * nf, 3)
    conv(p + "cas_offset_conv1", nf, 2 * nf, 3)
    conv(p + "cas_offset_conv2", nf, nf, 3)
    dcnpack(p + "cas_dcnpack")
    if with_tsa:
        p = "fusion."
        conv(p + "temporal_attn1", nf, nf, 3)
        conv(p + "temporal_attn2", nf, nf, 3)
        conv(p + "feat_fusion", nf, num_frame * nf, 1)
        conv(p + "spatial_attn1", nf, num_frame * nf, 1)
        conv(p + "spatial_attn2", nf, 2 * nf, 1)
        conv(p + "spatial_attn3", nf, nf, 3)
        conv(p + "spatial_attn4", nf, nf, 1)
        conv(p + "spatial_attn5", nf, nf, 3)
        conv(p + "spatial_attn_l1", nf, nf, 1)
        conv(p + "spatial_attn_l2", nf, 2 * nf, 3)
        conv(p + "spatial_attn_l3", nf, nf, 3)
        conv(p + "spatial_attn_add1", nf, nf, 1)
        conv(p + "spatial_attn_add2", nf, nf, 1)
    else:
        conv("fusion", nf, num_frame * nf, 1)
    for i in range(num_reconstruct_block):
        resblock(f"reconstruction.{i}")
    conv("upconv1", nf * 4, nf, 3)
    conv("upconv2", 64 * 4, nf, 3)
    conv("conv_hr", 64, 64, 3)
    conv("conv_last", 3, 64, 3)
    return sd
