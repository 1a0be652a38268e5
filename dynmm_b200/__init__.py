"""dynmm_b200 -- B200-native (sm_100a) implementation of DynMM's gated hot path.

``dynmm_b200.fusion``   drop-in ``SkipGateESANet`` / ``GlobalGate`` / ``DiffSoftmax``
                        (reference: FusionDynMM/src/models/model_skip_mod_globalgate.py)
``dynmm_b200.modality`` drop-in ``DynMMNet`` / ``DynMMNetV2`` / ``MMDL``
                        (reference: ModalityDynMM/{multimedia,affect}/*_dyn.py)
``dynmm_b200.ops``      tensor-level wrappers over the C ABI (include/dynmm_b200.h)
"""
__version__ = "0.1.0"
