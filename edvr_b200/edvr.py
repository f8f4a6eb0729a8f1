"""Drop-in module types for ``basicsr.models.archs.edvr_arch`` / ``arch_util`` (B3 boundary).

Same class names, constructor signatures, parameter names, shapes and initialisation as the reference
(/root/reference/basicsr/models/archs/edvr_arch.py:9-117 PCDAlignment, :120-214 TSAFusion, :217-269
PredeblurModule, :272-420 EDVR; arch_util.py:51-64 make_layer, :67-95 ResidualBlockNoBN), so reference
checkpoints load with ``strict=True``.  The modules only OWN parameters; ``forward`` hands them to the
B200 executor (edvr_b200/engine.py).  Packed weights are cached and re-packed when a parameter changes.

Inference (torch.no_grad) runs entirely on the sm_100a kernels through the fused executor.  With autograd enabled
(training, BASELINE cfg 5) the same graph is built from autograd Functions over the same kernels (edvr_b200/train.py:
tcgen05 forward / dgrad / wgrad for every convolution, our DCN forward + backward) on NHWC 16-bit activations
(`train_dtype`, bf16 by default) with fp32 master parameters.  There is no cuDNN / eager-PyTorch convolution on any path.
"""
import logging

import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init

from . import ops
from . import train as T
from .dcn import DCNv2Pack
from .engine import EDVREngine, _Arena, pack_pcd, pack_tsa, run_pcd, run_tsa


@torch.no_grad()
def default_init_weights(module_list, scale=1, bias_fill=0, **kwargs):
    if not isinstance(module_list, list):
        module_list = [module_list]
    for module in module_list:
        for m in module.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                init.kaiming_normal_(m.weight, **kwargs)
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.fill_(bias_fill)


def make_layer(basic_block, num_basic_block, **kwarg):
    return nn.Sequential(*[basic_block(**kwarg) for _ in range(num_basic_block)])


def _params_key(module):
    return tuple((p.data_ptr(), int(p._version)) for p in module.parameters())


def _wants_autograd(module, x):
    """True when the call must stay differentiable (training, BASELINE cfg 5).  CPU tensors are an error: like the
    reference's dcn (deform_conv.py:133-134) this package has no CPU implementation."""
    if not x.is_cuda:
        raise NotImplementedError(f"edvr_b200.{module.__class__.__name__} runs on CUDA tensors only")
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters()))


TRAIN_DTYPE = torch.bfloat16        # activations / gradients of the training graph (BASELINE cfg 5: bf16); fp16 selectable


def _to_nhwc(t):
    """fp32 NCHW (the reference modules' interface) -> NHWC 16-bit tensor of the training graph."""
    return t.permute(0, 2, 3, 1).to(TRAIN_DTYPE).contiguous()


def _from_nhwc(t, like):
    return t.permute(0, 3, 1, 2).to(like.dtype).contiguous()


def _sd(module, prefix=""):
    return {prefix + k: v for k, v in module.state_dict().items()}


class ResidualBlockNoBN(nn.Module):
    """x + conv2(relu(conv1(x))) * res_scale  (arch_util.py:67-95)."""

    def __init__(self, num_feat=64, res_scale=1, pytorch_init=False):
        super().__init__()
        self.res_scale = res_scale
        self.conv1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=True)
        self.relu = nn.ReLU(inplace=True)
        if not pytorch_init:
            default_init_weights([self.conv1, self.conv2], 0.1)
        self._cache = None

    def forward(self, x):
        C = self.conv1.in_channels
        if _wants_autograd(self, x):
            if self.res_scale != 1:
                raise ValueError("edvr_b200.ResidualBlockNoBN: res_scale != 1 is not covered")
            return _from_nhwc(T.resblock(self, _to_nhwc(x)), x)
        if C % 64 or self.res_scale != 1:
            raise ValueError(f"edvr_b200.ResidualBlockNoBN: num_feat={C} (multiple of 64) and res_scale={self.res_scale} (1) "
                             "are outside the tensor-core path; there is no cuDNN fallback")
        key = _params_key(self)
        if self._cache is None or self._cache[0] != key:
            self._cache = (key, ops.pack_conv(self.conv1.weight.detach().float(), self.conv1.bias.detach().float()),
                           ops.pack_conv(self.conv2.weight.detach().float(), self.conv2.bias.detach().float()))
        xv = ops.nchw_to_nhwc(x.float())
        t, o = ops.new_act(xv.N, xv.H, xv.W, C), ops.new_act(xv.N, xv.H, xv.W, C)
        ops.conv2d(self._cache[1], [xv], out16=t, act=ops.ACT_RELU)
        ops.conv2d(self._cache[2], [t], out16=o, act=ops.ACT_NONE, res16=xv)
        return ops.nhwc_to_nchw(o).to(x.dtype)


class PCDAlignment(nn.Module):
    """Pyramid, cascading and deformable alignment (edvr_arch.py:9-117)."""

    def __init__(self, num_feat=64, deformable_groups=8):
        super().__init__()
        self.offset_conv1 = nn.ModuleDict()
        self.offset_conv2 = nn.ModuleDict()
        self.offset_conv3 = nn.ModuleDict()
        self.dcn_pack = nn.ModuleDict()
        self.feat_conv = nn.ModuleDict()
        for i in range(3, 0, -1):
            level = f"l{i}"
            self.offset_conv1[level] = nn.Conv2d(num_feat * 2, num_feat, 3, 1, 1)
            self.offset_conv2[level] = nn.Conv2d(num_feat if i == 3 else num_feat * 2, num_feat, 3, 1, 1)
            if i < 3:
                self.offset_conv3[level] = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
                self.feat_conv[level] = nn.Conv2d(num_feat * 2, num_feat, 3, 1, 1)
            self.dcn_pack[level] = DCNv2Pack(num_feat, num_feat, 3, padding=1, deformable_groups=deformable_groups)
        self.cas_offset_conv1 = nn.Conv2d(num_feat * 2, num_feat, 3, 1, 1)
        self.cas_offset_conv2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.cas_dcnpack = DCNv2Pack(num_feat, num_feat, 3, padding=1, deformable_groups=deformable_groups)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        self.num_feat, self.deformable_groups = num_feat, deformable_groups
        self._cache = None

    def _autograd_forward(self, nbr, ref):
        return _from_nhwc(T.pcd_align(self, [_to_nhwc(t) for t in nbr], [_to_nhwc(t) for t in ref]), nbr[0])

    def forward(self, nbr_feat_l, ref_feat_l):
        x0 = nbr_feat_l[0]
        if _wants_autograd(self, x0):
            return self._autograd_forward(nbr_feat_l, ref_feat_l)
        if self.num_feat % 64:
            raise ValueError(f"edvr_b200.PCDAlignment: num_feat={self.num_feat} must be a multiple of 64; no cuDNN fallback")
        key = _params_key(self)
        if self._cache is None or self._cache[0] != key:
            p = {}
            pack_pcd(p, {k: v.detach().float() for k, v in _sd(self, "pcd.").items()}, "pcd.", self.deformable_groups)
            self._cache = (key, p, _Arena(x0.device))
        _, p, arena = self._cache
        nbr = [ops.nchw_to_nhwc(t.float()) for t in nbr_feat_l]
        ref = [ops.nchw_to_nhwc(t.float()) for t in ref_feat_l]
        out = ops.new_act(nbr[0].N, nbr[0].H, nbr[0].W, self.num_feat)
        run_pcd(arena, p, "pcd.", self.deformable_groups, None, nbr, ref, None, out)
        return ops.nhwc_to_nchw(out).to(x0.dtype)


class TSAFusion(nn.Module):
    """Temporal-spatial attention fusion (edvr_arch.py:120-214)."""

    def __init__(self, num_feat=64, num_frame=5, center_frame_idx=2):
        super().__init__()
        self.center_frame_idx = center_frame_idx
        self.temporal_attn1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.temporal_attn2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.feat_fusion = nn.Conv2d(num_frame * num_feat, num_feat, 1, 1)
        self.max_pool = nn.MaxPool2d(3, stride=2, padding=1)
        self.avg_pool = nn.AvgPool2d(3, stride=2, padding=1)
        self.spatial_attn1 = nn.Conv2d(num_frame * num_feat, num_feat, 1)
        self.spatial_attn2 = nn.Conv2d(num_feat * 2, num_feat, 1)
        self.spatial_attn3 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.spatial_attn4 = nn.Conv2d(num_feat, num_feat, 1)
        self.spatial_attn5 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.spatial_attn_l1 = nn.Conv2d(num_feat, num_feat, 1)
        self.spatial_attn_l2 = nn.Conv2d(num_feat * 2, num_feat, 3, 1, 1)
        self.spatial_attn_l3 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.spatial_attn_add1 = nn.Conv2d(num_feat, num_feat, 1)
        self.spatial_attn_add2 = nn.Conv2d(num_feat, num_feat, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.num_feat = num_feat
        self._cache = None

    def _autograd_forward(self, aligned):
        a = aligned.permute(0, 1, 3, 4, 2).to(TRAIN_DTYPE).contiguous()          # [b, t, h, w, c]
        return _from_nhwc(T.tsa_fusion(self, a), aligned)

    def forward(self, aligned_feat):
        b, t, c, h, w = aligned_feat.size()
        if _wants_autograd(self, aligned_feat):
            return self._autograd_forward(aligned_feat)
        if self.num_feat not in (64, 128, 256) or h % 4 or w % 4:
            raise ValueError(f"edvr_b200.TSAFusion: num_feat={self.num_feat} (64/128/256) and h, w = {h}, {w} (multiples of 4) "
                             "are outside the tensor-core path; there is no cuDNN fallback")
        key = _params_key(self)
        if self._cache is None or self._cache[0] != key:
            p = {}
            pack_tsa(p, {k: v.detach().float() for k, v in _sd(self, "tsa.").items()}, "tsa.")
            self._cache = (key, p, _Arena(aligned_feat.device))
        _, p, arena = self._cache
        av = ops.nchw_to_nhwc(aligned_feat.reshape(b * t, c, h, w).float())
        f16 = ops.new_act(b, h, w, c)
        run_tsa(arena, p, "tsa.", av, b, t, self.center_frame_idx, f16, None)
        return ops.nhwc_to_nchw(f16).to(aligned_feat.dtype)


class PredeblurModule(nn.Module):
    """Pre-deblur pyramid (edvr_arch.py:217-269)."""

    def __init__(self, num_in_ch=3, num_feat=64, hr_in=False):
        super().__init__()
        self.hr_in = hr_in
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        if self.hr_in:
            self.stride_conv_hr1 = nn.Conv2d(num_feat, num_feat, 3, 2, 1)
            self.stride_conv_hr2 = nn.Conv2d(num_feat, num_feat, 3, 2, 1)
        self.stride_conv_l2 = nn.Conv2d(num_feat, num_feat, 3, 2, 1)
        self.stride_conv_l3 = nn.Conv2d(num_feat, num_feat, 3, 2, 1)
        self.resblock_l3 = ResidualBlockNoBN(num_feat=num_feat)
        self.resblock_l2_1 = ResidualBlockNoBN(num_feat=num_feat)
        self.resblock_l2_2 = ResidualBlockNoBN(num_feat=num_feat)
        self.resblock_l1 = nn.ModuleList([ResidualBlockNoBN(num_feat=num_feat) for _ in range(5)])
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, x):
        if not x.is_cuda:
            raise NotImplementedError("edvr_b200.PredeblurModule runs on CUDA tensors only")
        return _from_nhwc(T.predeblur(self, _to_nhwc(x)), x)       # autograd Functions over the tensor-core kernels


class EDVR(nn.Module):
    """EDVR network, x4 video SR / restoration (edvr_arch.py:272-420)."""

    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_frame=5, deformable_groups=8,
                 num_extract_block=5, num_reconstruct_block=10, center_frame_idx=2, hr_in=False,
                 with_predeblur=False, with_tsa=True):
        super().__init__()
        self.center_frame_idx = num_frame // 2 if center_frame_idx is None else center_frame_idx
        self.hr_in, self.with_predeblur, self.with_tsa = hr_in, with_predeblur, with_tsa
        self.num_frame, self.num_feat = num_frame, num_feat
        if self.with_predeblur:
            self.predeblur = PredeblurModule(num_feat=num_feat, hr_in=self.hr_in)
            self.conv_1x1 = nn.Conv2d(num_feat, num_feat, 1, 1)
        else:
            self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.feature_extraction = make_layer(ResidualBlockNoBN, num_extract_block, num_feat=num_feat)
        self.conv_l2_1 = nn.Conv2d(num_feat, num_feat, 3, 2, 1)
        self.conv_l2_2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_l3_1 = nn.Conv2d(num_feat, num_feat, 3, 2, 1)
        self.conv_l3_2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.pcd_align = PCDAlignment(num_feat=num_feat, deformable_groups=deformable_groups)
        if self.with_tsa:
            self.fusion = TSAFusion(num_feat=num_feat, num_frame=num_frame, center_frame_idx=self.center_frame_idx)
        else:
            self.fusion = nn.Conv2d(num_frame * num_feat, num_feat, 1, 1)
        self.reconstruction = make_layer(ResidualBlockNoBN, num_reconstruct_block, num_feat=num_feat)
        self.upconv1 = nn.Conv2d(num_feat, num_feat * 4, 3, 1, 1)
        self.upconv2 = nn.Conv2d(num_feat, 64 * 4, 3, 1, 1)
        self.pixel_shuffle = nn.PixelShuffle(2)
        self.conv_hr = nn.Conv2d(64, 64, 3, 1, 1)
        self.conv_last = nn.Conv2d(64, 3, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        self._engine = None

    def engine(self):
        """The B200 executor for the current parameters (re-packed when any parameter changed)."""
        key = _params_key(self)
        if self._engine is None or self._engine[0] != key:
            dev = next(self.parameters()).device
            self._engine = (key, EDVREngine(self.state_dict(), self.num_frame, self.center_frame_idx,
                                            self.hr_in, device=dev))
        return self._engine[1]

    def _autograd_forward(self, x):
        return T.edvr_forward(self, x.float(), dtype=getattr(self, "train_dtype", TRAIN_DTYPE)).to(x.dtype)

    def forward(self, x):
        b, t, c, h, w = x.size()
        if self.hr_in:
            assert h % 16 == 0 and w % 16 == 0, "The height and width must be multiple of 16."
        else:
            assert h % 4 == 0 and w % 4 == 0, "The height and width must be multiple of 4."
        if not x.is_cuda:
            # no CPU path (north_star); the reference cannot run EDVR on CPU either: its dcn raises at deform_conv.py:133-134
            raise NotImplementedError("edvr_b200.EDVR runs on CUDA tensors only")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._autograd_forward(x)       # training (BASELINE cfg 5): differentiable graph, see train.py
        return self.engine().forward(x.float()).to(x.dtype)

    @torch.no_grad()
    def forward_video(self, frames, clips_per_step=4, padding="reflection_circle"):
        """Restore every frame of one sequence [F, 3, h, w] with sliding windows (padding modes of the reference datasets);
        per-frame features are shared between windows - see EDVREngine.forward_video."""
        return self.engine().forward_video(frames.float(), clips_per_step, padding).to(frames.dtype)


def load_network(net, load_path, strict=True, param_key="params"):
    """Checkpoint wire format of the reference (`BaseModel.load_network`, basicsr/models/base_model.py:238-262; save side
    `save_network` :171-201): a `torch.save`d dict whose `param_key` entry ('params'; None = the file IS the state_dict)
    holds the state_dict, possibly with DataParallel's 'module.' prefix on every key.  Same behaviour: key differences are
    logged, with strict=False tensors of a different size are skipped, then `load_state_dict(strict=strict)`."""
    if isinstance(net, (nn.DataParallel, nn.parallel.DistributedDataParallel)):
        net = net.module
    log = logging.getLogger("basicsr")
    log.info(f"Loading {net.__class__.__name__} model from {load_path}.")
    load_net = torch.load(load_path, map_location=lambda storage, loc: storage)
    if param_key is not None:
        load_net = load_net[param_key]
    load_net = {(k[7:] if k.startswith("module.") else k): v for k, v in load_net.items()}
    cur = net.state_dict()
    cur_keys, new_keys = set(cur.keys()), set(load_net.keys())
    if cur_keys != new_keys:
        log.warning("Current net - loaded net:")
        for k in sorted(cur_keys - new_keys):
            log.warning(f"  {k}")
        log.warning("Loaded net - current net:")
        for k in sorted(new_keys - cur_keys):
            log.warning(f"  {k}")
    if not strict:
        for k in cur_keys & new_keys:
            if cur[k].size() != load_net[k].size():
                log.warning(f"Size different, ignore [{k}]: crt_net: {cur[k].shape}; load_net: {load_net[k].shape}")
                load_net[k + ".ignore"] = load_net.pop(k)
    net.load_state_dict(load_net, strict=strict)
    return net


def save_network(net, save_path, param_key="params"):
    """Write a checkpoint the reference's `load_network` reads (base_model.py:171-201: 'module.' stripped, CPU tensors)."""
    if isinstance(net, (nn.DataParallel, nn.parallel.DistributedDataParallel)):
        net = net.module
    sd = {(k[7:] if k.startswith("module.") else k): v.cpu() for k, v in net.state_dict().items()}
    torch.save({param_key: sd} if param_key is not None else sd, save_path)


# ---- official-release key map (SURVEY §8 f4 tail) -------------------------------------------------------------------------------
_OFFICIAL_RULES = (      # (regex on the BasicSR key, replacement) - first match wins; restates the renames of
    # /root/reference/scripts/model_conversion/convert_models.py:17-103 (EDVR official release -> BasicSR names)
    (r"^predeblur\.stride_conv_hr1\.", "pre_deblur.conv_first_2."),
    (r"^predeblur\.stride_conv_hr2\.", "pre_deblur.conv_first_3."),
    (r"^predeblur\.conv_first\.", "pre_deblur.conv_first_1."),
    (r"^predeblur\.stride_conv_l([23])\.", r"pre_deblur.deblur_L\1_conv."),
    (r"^predeblur\.resblock_l3\.", "pre_deblur.RB_L3_1."),
    (r"^predeblur\.resblock_l2_([12])\.", r"pre_deblur.RB_L2_\1."),
    (r"^predeblur\.resblock_l1\.(\d+)\.", lambda m: f"pre_deblur.RB_L1_{int(m.group(1)) + 1}."),
    (r"^conv_l([23])_([12])\.", r"fea_L\1_conv\2."),
    (r"^pcd_align\.dcn_pack\.l(\d)\.conv_offset\.", r"pcd_align.L\1_dcnpack.conv_offset_mask."),
    (r"^pcd_align\.dcn_pack\.l(\d)\.", r"pcd_align.L\1_dcnpack."),
    (r"^pcd_align\.offset_conv(\d)\.l(\d)\.", r"pcd_align.L\2_offset_conv\1."),
    (r"^pcd_align\.feat_conv\.l(\d)\.", r"pcd_align.L\1_fea_conv."),
    (r"^pcd_align\.cas_dcnpack\.conv_offset\.", "pcd_align.cas_dcnpack.conv_offset_mask."),
    (r"^fusion\.temporal_attn1\.", "tsa_fusion.tAtt_2."),
    (r"^fusion\.temporal_attn2\.", "tsa_fusion.tAtt_1."),
    (r"^fusion\.feat_fusion\.", "tsa_fusion.fea_fusion."),
    (r"^fusion\.spatial_attn_add(\d)\.", r"tsa_fusion.sAtt_add_\1."),
    (r"^fusion\.spatial_attn_l(\d)\.", r"tsa_fusion.sAtt_L\1."),
    (r"^fusion\.spatial_attn(\d)\.", r"tsa_fusion.sAtt_\1."),
    (r"^reconstruction\.", "recon_trunk."),
    (r"^conv_hr\.", "HRconv."),
    (r"^fusion\.", "tsa_fusion."),
)


def official_key(key):
    """Name of BasicSR parameter `key` in the original EDVR release's checkpoints."""
    import re
    for pat, rep in _OFFICIAL_RULES:
        new, n = re.subn(pat, rep, key, count=1)
        if n:
            return new
    return key          # conv_first, feature_extraction, cas_offset_conv*, upconv*, conv_last, conv_1x1 keep their names


def convert_official_state_dict(official_sd, net):
    """Official-release EDVR weights -> a state_dict for `net` (ours or the reference's EDVR): what
    scripts/model_conversion/convert_models.py::convert_edvr does with hard-coded paths."""
    return {k: official_sd[official_key(k)] for k in net.state_dict()}
