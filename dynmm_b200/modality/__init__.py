"""Modality-level DynMM (MM-IMDB, CMU-MOSEI) -- drop-in modules."""
from .affect import DynMMNet as DynMMNet3, DynMMNetV2, build_mosei_experts  # noqa: F401
from .common_models import MLP, Concat, Identity, Linear, MaxOut_MLP, Maxout, Sequential, Transformer  # noqa: F401
from .gating import DiffSoftmax, mix, routed_mix  # noqa: F401
from .imdb import DynMMNet  # noqa: F401
from .supervised import MMDL  # noqa: F401
