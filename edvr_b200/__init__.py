"""B200-native EDVR hot path.  Public names are resolved lazily so that `import edvr_b200` stays cheap:

    from edvr_b200 import EDVR, PCDAlignment, TSAFusion, ResidualBlockNoBN, PredeblurModule     # drop-in module types
    from edvr_b200 import DCNv2Pack, ModulatedDeformConvPack, modulated_deform_conv, deform_conv # dcn operator API
    from edvr_b200 import EDVREngine                                                           # fused inference executor
    from edvr_b200 import read_img_seq, tensor2img                                               # frame staging on the device
"""
_EXPORTS = {
    "EDVR": "edvr", "PCDAlignment": "edvr", "TSAFusion": "edvr", "ResidualBlockNoBN": "edvr", "PredeblurModule": "edvr",
    "load_network": "edvr", "save_network": "edvr", "convert_official_state_dict": "edvr",
    "DCNv2Pack": "dcn", "ModulatedDeformConv": "dcn", "ModulatedDeformConvPack": "dcn", "modulated_deform_conv": "dcn",
    "DeformConv": "dcn", "DeformConvPack": "dcn", "deform_conv": "dcn",
    "EDVREngine": "engine",
    "read_img_seq": "img", "tensor2img": "img", "frames_to_tensor": "img", "tensor_to_bytes": "img",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        import importlib
        return getattr(importlib.import_module("." + _EXPORTS[name], __name__), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
