"""Import-path shims so the reference's scripts and pickles resolve to this package.

``install_aliases()`` registers
  unimodals.common_models, fusions.common_fusions, training_structures.Supervised_Learning
(the MultiBench paths imported at imdb_dyn.py:10-13 / affect_dyn.py:12-15) and
  src.models.model_skip_mod_globalgate, src.models.model_skip_mod
(imported by FusionDynMM/src/build_model.py:11-12) in ``sys.modules``.
"""
from __future__ import annotations

import sys
import types


def install_aliases() -> None:
    from . import common_models, supervised
    from ..fusion import local_gate as fusion_local_gate
    from ..fusion import modules as fusion_modules

    def alias(name, module):
        parts = name.split(".")
        for i in range(1, len(parts)):
            pkg = ".".join(parts[:i])
            if pkg not in sys.modules:
                sys.modules[pkg] = types.ModuleType(pkg)
        sys.modules[name] = module
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], module)

    alias("unimodals.common_models", common_models)
    fusions = types.ModuleType("fusions.common_fusions")
    fusions.Concat = common_models.Concat
    alias("fusions.common_fusions", fusions)
    alias("training_structures.Supervised_Learning", supervised)
    alias("src.models.model_skip_mod_globalgate", fusion_modules)
    alias("src.models.model_skip_mod", fusion_local_gate)
