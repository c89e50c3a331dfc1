"""Import the UNMODIFIED reference modules read-only from /root/reference (this container only).

Test/fixture infrastructure: used by make_golden.py to produce the committed golden vectors under
tests/golden/.  Nothing in the product package, the `-m gpu` tests, smoke() or bench.py imports this
file -- /root/reference does not exist on the GPU box.

Recipe: stub `detectron2` / `fvcore` with the handful of symbols the hot-path
files use, and mount the reference packages as *namespace* modules so their heavy `__init__.py`
files (datasets, timm, Detectron2 data) are bypassed.
"""
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("DVIS_REFERENCE_ROOT", "/root/reference")
P = os.path.join(REF_ROOT, "DVIS_Plus")


def available() -> bool:
    return os.path.isdir(P)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _ns(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


class _Registry(dict):
    def __init__(self, name="registry"):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        def deco(o):
            self[o.__name__] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, name):
        return self[name]


def _configurable(init_func=None, *, from_config=None):
    # pass-through: the fixtures construct modules with explicit keyword arguments
    if init_func is not None:
        return init_func
    return lambda f: f


class _ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class _Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d: conv followed by optional norm and activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = super().forward(x)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def _get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    assert norm == "GN", norm
    return nn.GroupNorm(32, out_channels)


def _c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


_installed = False


def install():
    """Idempotently install the stubs + namespace packages into sys.modules."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {P}")
    # the compiled extension the reference op file insists on importing; CPU path never calls it
    if "MultiScaleDeformableAttention" not in sys.modules:
        def _no_ext(*a, **k):
            raise RuntimeError("Not implemented on the CPU")
        _mod("MultiScaleDeformableAttention", ms_deform_attn_forward=_no_ext, ms_deform_attn_backward=_no_ext)
    _mod("detectron2")
    _mod("detectron2.config", configurable=_configurable)
    _mod("detectron2.layers", Conv2d=_Conv2d, ShapeSpec=_ShapeSpec, get_norm=_get_norm)
    _mod("detectron2.modeling", SEM_SEG_HEADS_REGISTRY=_Registry("SEM_SEG_HEADS"))
    _mod("detectron2.utils")
    _mod("detectron2.utils.registry", Registry=_Registry)
    _mod("fvcore")
    _mod("fvcore.nn")
    _mod("fvcore.nn.weight_init", c2_xavier_fill=_c2_xavier_fill)
    sys.modules["fvcore.nn"].weight_init = sys.modules["fvcore.nn.weight_init"]
    _ns("mask2former", f"{P}/mask2former")
    _ns("mask2former.modeling", f"{P}/mask2former/modeling")
    _ns("mask2former.modeling.pixel_decoder", f"{P}/mask2former/modeling/pixel_decoder")
    _ns("mask2former.modeling.transformer_decoder", f"{P}/mask2former/modeling/transformer_decoder")
    _ns("mask2former_video", f"{P}/mask2former_video")
    _ns("mask2former_video.modeling", f"{P}/mask2former_video/modeling")
    _ns("mask2former_video.modeling.transformer_decoder", f"{P}/mask2former_video/modeling/transformer_decoder")
    _ns("dvis_Plus", f"{P}/dvis_Plus")
    _ns("dvis_daq", f"{REF_ROOT}/DVIS_DAQ/dvis_daq")
    _installed = True


def load():
    """Return a namespace with the reference classes / functions on the hot path."""
    install()
    import importlib
    ops_func = importlib.import_module("mask2former.modeling.pixel_decoder.ops.functions.ms_deform_attn_func")
    ops_mod = importlib.import_module("mask2former.modeling.pixel_decoder.ops.modules.ms_deform_attn")
    pix = importlib.import_module("mask2former.modeling.pixel_decoder.msdeformattn")
    dec = importlib.import_module("dvis_Plus.video_mask2former_transformer_decoder")
    trk = importlib.import_module("dvis_Plus.tracker")
    rfn = importlib.import_module("dvis_Plus.refiner")
    daq_trk = importlib.import_module("dvis_daq.track_module")
    daq_rfn = importlib.import_module("dvis_daq.refiner")
    daq_slot = importlib.import_module("dvis_daq.slot_attention")
    return types.SimpleNamespace(
        VideoInstanceCutter=daq_trk.VideoInstanceCutter,
        DAQTemporalRefiner=daq_rfn.TemporalRefiner,
        SlotCrossAttentionLayer=daq_slot.SlotCrossAttentionLayer,
        ms_deform_attn_core_pytorch=ops_func.ms_deform_attn_core_pytorch,
        MSDeformAttn=ops_mod.MSDeformAttn,
        MSDeformAttnPixelDecoder=pix.MSDeformAttnPixelDecoder,
        MSDeformAttnTransformerEncoderLayer=pix.MSDeformAttnTransformerEncoderLayer,
        Decoder_dvisPlus=dec.VideoMultiScaleMaskedTransformerDecoder_dvisPlus,
        ReferringTracker_noiser=trk.ReferringTracker_noiser,
        TemporalRefiner=rfn.TemporalRefiner,
        ShapeSpec=_ShapeSpec,
    )


def load_meta_architecture():
    """Import P/dvis_Plus/meta_architecture.py (the post-processing methods inference_video_vis / vps / vss and
    post_processing live on the meta-architecture classes, py:255-301,758-979).  The file's imports of Detectron2
    structures, the criterion and the matcher are stubbed: none of them is touched by those methods."""
    install()
    import importlib
    _mod("detectron2.data", MetadataCatalog=types.SimpleNamespace(get=lambda name: types.SimpleNamespace()))
    sys.modules["detectron2.modeling"].__dict__.update(
        META_ARCH_REGISTRY=_Registry("META_ARCH"), build_backbone=None, build_sem_seg_head=None)
    _mod("detectron2.modeling.backbone", Backbone=nn.Module)
    _mod("detectron2.structures", Boxes=object, ImageList=object, Instances=object, BitMasks=object)
    _mod("mask2former_video.modeling.criterion", VideoSetCriterion=object)
    _mod("mask2former_video.modeling.matcher", VideoHungarianMatcher=object, VideoHungarianMatcher_Consistent=object)
    _ns("mask2former_video.utils", f"{P}/mask2former_video/utils")
    return importlib.import_module("dvis_Plus.meta_architecture")


def load_daq_meta_architecture():
    """Import D/dvis_daq/meta_architecture.py (DVIS_DAQ_online.run_window_inference, py:488-597) with the same kind of
    stubs: pycocotools and the DAQ criterion / matcher are never touched by the inference window loop."""
    load_meta_architecture()
    import importlib
    _mod("pycocotools")
    _mod("pycocotools.mask")
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    _mod("dvis_daq.matcher", FrameMatcher=object, NewInsHungarianMatcher=object)
    _mod("dvis_daq.criterion", DAQCriterion=object)
    return importlib.import_module("dvis_daq.meta_architecture")
