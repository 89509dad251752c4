"""Import the UNMODIFIED reference (arcee-ai/DALM at /root/reference) in the build container.

TEST INFRASTRUCTURE ONLY. Used by oracle/make_golden.py to generate tests/golden/* and by tests that are skipped when
/root/reference is absent (it never exists on the GPU box). `peft` and `accelerate` are not installed here; they are
(and `hnswlib`, for dalm/eval/utils.py) replaced by MagicMock stubs AFTER torch/transformers are imported (SURVEY §8c recipe), which is enough for the loss
functions, wrappers' forward/mean_pooling and the batch builders to run as the reference wrote them.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("DALM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dalm"))


def load() -> SimpleNamespace:
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch  # noqa: F401
    import transformers  # noqa: F401  (must be imported before the stubs are installed)

    saved_dalm = {k: v for k, v in sys.modules.items() if k == "dalm" or k.startswith("dalm.")}
    for k in saved_dalm:
        del sys.modules[k]
    stubbed = []
    for name in ("peft", "accelerate", "accelerate.logging", "accelerate.utils", "hnswlib"):
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
            stubbed.append(name)
    # the repo's own `dalm` alias package installs a meta-path finder mapping dalm.* -> dalm_b200.*: park it
    parked = [f for f in sys.meta_path if getattr(f, "__name__", "") == "_LazyAlias"]
    for f in parked:
        sys.meta_path.remove(f)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from dalm.training.utils import train_utils
        from dalm.models.rag_e2e_base_model import AutoModelForRagE2E
        from dalm.models.retriever_only_base_model import AutoModelForSentenceEmbedding
        from dalm.training.utils.rag_e2e_dataloader_utils import preprocess_dataset as preprocess_e2e
        from dalm.training.utils.retriever_only_dataloader_utils import preprocess_dataset as preprocess_retriever
        from dalm.utils import eos_mask
        from dalm.eval import utils as eval_utils
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for f in parked:
            sys.meta_path.insert(0, f)
        ref_modules = {k: v for k, v in sys.modules.items() if k == "dalm" or k.startswith("dalm.")}
        for k in ref_modules:
            del sys.modules[k]
        sys.modules.update(saved_dalm)
        for name in stubbed:            # the loaded reference modules keep their references; nobody else should see the mocks
            sys.modules.pop(name, None)  # (transformers probes importlib.util.find_spec("peft") and chokes on a MagicMock)
    return SimpleNamespace(
        train_utils=train_utils, AutoModelForRagE2E=AutoModelForRagE2E,
        AutoModelForSentenceEmbedding=AutoModelForSentenceEmbedding, preprocess_e2e=preprocess_e2e,
        preprocess_retriever=preprocess_retriever, eos_mask=eos_mask, eval_utils=eval_utils,
    )
