"""Host-side mirrors of the reference's hot-path modules (same class names, constructor arguments, parameter
names and output dictionaries), running on the libdvis_b200 kernels."""
from .ms_deform_attn import MSDeformAttn, MSDeformAttnFunction  # noqa: F401
from .pixel_decoder import MSDeformAttnPixelDecoder  # noqa: F401
from .predictor import VideoMultiScaleMaskedTransformerDecoder_dvisPlus  # noqa: F401
from .tracker import ReferringTracker_noiser  # noqa: F401
from .refiner import TemporalRefiner  # noqa: F401
from .daq import SlotCrossAttentionLayer, VideoInstanceCutter  # noqa: F401
from .daq import TemporalRefiner as DAQTemporalRefiner  # noqa: F401
