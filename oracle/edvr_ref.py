"""oracle/edvr_ref.py — TEST INFRASTRUCTURE ONLY (the checker, never the product).

Functional fp32 restatement of the reference EDVR forward graph in plain PyTorch ops,
driven by a reference-format ``state_dict`` (same keys as
``basicsr.models.archs.edvr_arch.EDVR``).  It exists because /root/reference cannot
travel to the GPU box: this port is validated HERE, bit-for-bit on CPU, against the
imported reference (tests/test_oracle.py::test_edvr_port_matches_reference_import and
oracle/make_golden.py) and then stands in for it there.

Follows (paths relative to /root/reference/basicsr/models/archs/):
  EDVR.forward ........... edvr_arch.py:358-420
  PCDAlignment.forward ... edvr_arch.py:76-117
  TSAFusion.forward ...... edvr_arch.py:161-214
  PredeblurModule ........ edvr_arch.py:250-269
  DCNv2Pack.forward ...... arch_util.py:243-257
  ResidualBlockNoBN ...... arch_util.py:92-95

``dcn`` is the modulated deformable conv to use, with the reference B2 signature
``dcn(x, offset, mask, weight, bias, stride, padding, dilation, groups, dg)``; the
default is torchvision's CPU/GPU ``deform_conv2d`` (identical forward semantics, see
SURVEY §8c).
"""
import torch
import torch.nn.functional as F


def dcn_torchvision(x, offset, mask, weight, bias, stride, padding, dilation, groups, dg):
    from torchvision.ops import deform_conv2d
    return deform_conv2d(x, offset, weight, bias, stride=stride, padding=padding,
                         dilation=dilation, mask=mask)


def _conv(sd, key, x, stride=1, padding=1):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride, padding)


def _lrelu(x):
    return F.leaky_relu(x, 0.1)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def _resblock(sd, key, x):
    return x + _conv(sd, key + ".conv2", F.relu(_conv(sd, key + ".conv1", x)))


def _dcn_pack(sd, key, x, feat, dg, dcn):
    raw = _conv(sd, key + ".conv_offset", feat)
    n_off = raw.shape[1] // 3 * 2           # chunk(3)+cat(o1,o2) == first two thirds
    offset, mask = raw[:, :n_off], torch.sigmoid(raw[:, n_off:])
    return dcn(x, offset.contiguous(), mask.contiguous(), sd[key + ".weight"],
               sd.get(key + ".bias"), 1, 1, 1, 1, dg)


def pcd_align(sd, nbr, ref, dg, dcn, prefix="pcd_align"):
    """nbr/ref: lists [L1, L2, L3] of (b, c, h, w)."""
    p = prefix + "."
    up_off = up_feat = None
    for lvl in (3, 2, 1):
        L = f"l{lvl}"
        off = _lrelu(_conv(sd, p + "offset_conv1." + L, torch.cat([nbr[lvl - 1], ref[lvl - 1]], 1)))
        if lvl == 3:
            off = _lrelu(_conv(sd, p + "offset_conv2." + L, off))
        else:
            off = _lrelu(_conv(sd, p + "offset_conv2." + L, torch.cat([off, up_off], 1)))
            off = _lrelu(_conv(sd, p + "offset_conv3." + L, off))
        feat = _dcn_pack(sd, p + "dcn_pack." + L, nbr[lvl - 1], off, dg, dcn)
        if lvl < 3:
            feat = _conv(sd, p + "feat_conv." + L, torch.cat([feat, up_feat], 1))
        if lvl > 1:
            feat = _lrelu(feat)
            up_off = _up2(off) * 2
            up_feat = _up2(feat)
    off = torch.cat([feat, ref[0]], 1)
    off = _lrelu(_conv(sd, p + "cas_offset_conv2", _lrelu(_conv(sd, p + "cas_offset_conv1", off))))
    return _lrelu(_dcn_pack(sd, p + "cas_dcnpack", feat, off, dg, dcn))


def tsa_fusion(sd, aligned, center, prefix="fusion"):
    """aligned: (b, t, c, h, w) -> (b, c, h, w)."""
    p = prefix + "."
    b, t, c, h, w = aligned.shape
    emb_ref = _conv(sd, p + "temporal_attn1", aligned[:, center])
    emb = _conv(sd, p + "temporal_attn2", aligned.reshape(-1, c, h, w)).view(b, t, -1, h, w)
    prob = torch.sigmoid((emb * emb_ref.unsqueeze(1)).sum(2))             # (b, t, h, w)
    x = (aligned * prob.unsqueeze(2)).reshape(b, t * c, h, w)
    feat = _lrelu(_conv(sd, p + "feat_fusion", x, padding=0))
    attn = _lrelu(_conv(sd, p + "spatial_attn1", x, padding=0))
    pooled = torch.cat([F.max_pool2d(attn, 3, 2, 1), F.avg_pool2d(attn, 3, 2, 1)], 1)
    attn = _lrelu(_conv(sd, p + "spatial_attn2", pooled, padding=0))
    lvl = _lrelu(_conv(sd, p + "spatial_attn_l1", attn, padding=0))
    pooled = torch.cat([F.max_pool2d(lvl, 3, 2, 1), F.avg_pool2d(lvl, 3, 2, 1)], 1)
    lvl = _lrelu(_conv(sd, p + "spatial_attn_l2", pooled))
    lvl = _up2(_lrelu(_conv(sd, p + "spatial_attn_l3", lvl)))
    attn = _lrelu(_conv(sd, p + "spatial_attn3", attn)) + lvl
    attn = _up2(_lrelu(_conv(sd, p + "spatial_attn4", attn, padding=0)))
    attn = _conv(sd, p + "spatial_attn5", attn)
    add = _conv(sd, p + "spatial_attn_add2",
                _lrelu(_conv(sd, p + "spatial_attn_add1", attn, padding=0)), padding=0)
    return feat * torch.sigmoid(attn) * 2 + add


def predeblur(sd, x, hr_in, prefix="predeblur"):
    p = prefix + "."
    l1 = _lrelu(_conv(sd, p + "conv_first", x))
    if hr_in:
        l1 = _lrelu(_conv(sd, p + "stride_conv_hr1", l1, stride=2))
        l1 = _lrelu(_conv(sd, p + "stride_conv_hr2", l1, stride=2))
    l2 = _lrelu(_conv(sd, p + "stride_conv_l2", l1, stride=2))
    l3 = _lrelu(_conv(sd, p + "stride_conv_l3", l2, stride=2))
    l3 = _up2(_resblock(sd, p + "resblock_l3", l3))
    l2 = _resblock(sd, p + "resblock_l2_1", l2) + l3
    l2 = _up2(_resblock(sd, p + "resblock_l2_2", l2))
    for i in range(2):
        l1 = _resblock(sd, p + f"resblock_l1.{i}", l1)
    l1 = l1 + l2
    for i in range(2, 5):
        l1 = _resblock(sd, p + f"resblock_l1.{i}", l1)
    return l1


def config_from_state_dict(sd, num_frame=None, hr_in=False):
    """Recover the constructor kwargs the state_dict implies (num_frame needs TSA or a hint)."""
    nf = sd["conv_l2_1.weight"].shape[0]
    cfg = dict(num_feat=nf,
               deformable_groups=sd["pcd_align.dcn_pack.l1.conv_offset.weight"].shape[0] // 27,
               num_extract_block=len({k.split(".")[1] for k in sd if k.startswith("feature_extraction.")}),
               num_reconstruct_block=len({k.split(".")[1] for k in sd if k.startswith("reconstruction.")}),
               with_predeblur=any(k.startswith("predeblur.") for k in sd),
               with_tsa="fusion.feat_fusion.weight" in sd, hr_in=hr_in)
    fkey = "fusion.feat_fusion.weight" if cfg["with_tsa"] else "fusion.weight"
    cfg["num_frame"] = num_frame or sd[fkey].shape[1] // nf
    return cfg


@torch.no_grad()
def edvr_forward(sd, x, center_frame_idx=None, hr_in=False, dcn=dcn_torchvision,
                 return_intermediates=False):
    """x: (b, t, 3, h, w) fp32 -> (b, 3, 4h, 4w)  (or (b,3,h,w) if hr_in)."""
    cfg = config_from_state_dict(sd, num_frame=x.shape[1], hr_in=hr_in)
    b, t, c, h, w = x.shape
    center = t // 2 if center_frame_idx is None else center_frame_idx
    dg = cfg["deformable_groups"]
    inter = {}
    x_center = x[:, center].contiguous()
    if cfg["with_predeblur"]:
        l1 = _conv(sd, "conv_1x1", predeblur(sd, x.view(-1, c, h, w), hr_in), padding=0)
        if hr_in:
            h, w = h // 4, w // 4
    else:
        l1 = _lrelu(_conv(sd, "conv_first", x.view(-1, c, h, w)))
    for i in range(cfg["num_extract_block"]):
        l1 = _resblock(sd, f"feature_extraction.{i}", l1)
    l2 = _lrelu(_conv(sd, "conv_l2_2", _lrelu(_conv(sd, "conv_l2_1", l1, stride=2))))
    l3 = _lrelu(_conv(sd, "conv_l3_2", _lrelu(_conv(sd, "conv_l3_1", l2, stride=2))))
    l1 = l1.view(b, t, -1, h, w)
    l2 = l2.view(b, t, -1, h // 2, w // 2)
    l3 = l3.view(b, t, -1, h // 4, w // 4)
    inter["feat_l1"], inter["feat_l2"], inter["feat_l3"] = l1, l2, l3
    ref = [l1[:, center], l2[:, center], l3[:, center]]
    aligned = torch.stack([pcd_align(sd, [l1[:, i], l2[:, i], l3[:, i]], ref, dg, dcn)
                           for i in range(t)], 1)
    inter["aligned"] = aligned
    if cfg["with_tsa"]:
        feat = tsa_fusion(sd, aligned, center)
    else:
        feat = _conv(sd, "fusion", aligned.view(b, -1, h, w), padding=0)
    inter["fused"] = feat
    out = feat
    for i in range(cfg["num_reconstruct_block"]):
        out = _resblock(sd, f"reconstruction.{i}", out)
    inter["trunk"] = out
    out = _lrelu(F.pixel_shuffle(_conv(sd, "upconv1", out), 2))
    out = _lrelu(F.pixel_shuffle(_conv(sd, "upconv2", out), 2))
    out = _conv(sd, "conv_last", _lrelu(_conv(sd, "conv_hr", out)))
    base = x_center if hr_in else F.interpolate(x_center, scale_factor=4, mode="bilinear",
                                                align_corners=False)
    out = out + base
    return (out, inter) if return_intermediates else out


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
