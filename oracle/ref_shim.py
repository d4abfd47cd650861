"""Import shim that makes the UNMODIFIED reference importable (TEST INFRASTRUCTURE ONLY).

Used by ``oracle/gen_golden.py`` in the build container (where ``/root/reference`` exists), and — through the
package installed into ``oracle/_ref`` by ``oracle/install_ref.py`` — by ``bench.py --impl reference`` /
``cpu_baseline`` and ``tests/test_gpu_trainer_dropin.py`` on the GPU box, where ``/root/reference`` does not exist.
The product package never imports this module.

What it does (nothing is copied from the reference):
* puts ``oracle/standin`` on ``sys.path`` so ``import torch_scatter`` resolves to
  the stand-in (``oracle/standin/torch_scatter/__init__.py``);
* installs empty stub modules for the two logging-only dependencies that are
  missing from this image: ``rdl_ml_utils.handlers.wandb_handler.WanDBHandler``
  (reference ``relgat_projector/utils/logging_adapter.py:6``) and
  ``plwordnet_ml.embedder.constants.wandb.WandbConfig`` (reference
  ``relgat_projector/base/constants.py:33``);
* puts the reference root on ``sys.path``.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """Where the reference package can be imported from: $RELGAT_REFERENCE_ROOT, the read-only source tree in the build
    container, or the copy pip-installed into oracle/_ref (the only one present on the GPU box); None if none is."""
    cands = [os.environ.get("RELGAT_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "relgat_projector", "core", "model", "layer.py")):
            return c
    return None


REFERENCE_ROOT = reference_root()


def reference_available() -> bool:
    return reference_root() is not None


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def import_reference():
    """Returns the imported ``relgat_projector`` package of the reference."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not found (neither /root/reference nor oracle/_ref: run oracle/install_ref.py)")
    standin = os.path.join(_HERE, "standin")
    if standin not in sys.path:
        sys.path.insert(0, standin)
    if root not in sys.path:
        sys.path.insert(0, root)

    class WanDBHandler:  # logging sink, never reached with log_to_wandb=False
        @staticmethod
        def init_wandb(*a, **k):
            return None

        @staticmethod
        def log_metrics(*a, **k):
            return None

        @staticmethod
        def finish_wand(*a, **k):
            return None

        @staticmethod
        def finish_wandb(*a, **k):
            return None

    class WandbConfig:
        PROJECT_NAME = "stub"
        PROJECT_TAGS = []
        PREFIX_RUN = "run_"
        BASE_RUN_NAME = "stub"

    _stub("rdl_ml_utils")
    _stub("rdl_ml_utils.handlers")
    _stub("rdl_ml_utils.handlers.wandb_handler", WanDBHandler=WanDBHandler)
    _stub("plwordnet_ml")
    _stub("plwordnet_ml.embedder")
    _stub("plwordnet_ml.embedder.constants")
    _stub("plwordnet_ml.embedder.constants.wandb", WandbConfig=WandbConfig)

    import relgat_projector  # noqa: F401  (the reference package)

    return relgat_projector
