"""Fusion-level DynMM (RGB-D ESANet with a global gate) -- drop-in modules."""
from .modules import (BasicBlock, Bottleneck, ConvBNAct, Decoder, DecoderModule, DiffSoftmax, GlobalGate,  # noqa: F401
                      NonBottleneck1D, PyramidPoolingModule, ResNet, ResNet18, ResNet34, ResNet50, SkipGateESANet,
                      SqueezeAndExcitation, SqueezeAndExciteFusionAdd, Upsample, get_context_module)
from .local_gate import SkipESANet, SqueezeAndExcitationWeight, SqueezeAndExciteReweigh  # noqa: F401,E402
from .build import build_model  # noqa: F401,E402
from .pipeline import EvalPipeline  # noqa: F401,E402
from .loss import CrossEntropyLoss2d  # noqa: F401,E402
