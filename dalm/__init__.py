"""Alias package: `import dalm...` resolves to dalm_b200's drop-in modules, so code written against the reference's
import paths (dalm.models.*, dalm.training.*, dalm.cli, dalm.utils) runs unchanged on the B200 build."""
import importlib
import sys

import dalm_b200
from dalm_b200 import __version__  # noqa: F401

_ALIASES = [
    "cli", "utils", "models", "models.rag_e2e_base_model", "models.retriever_only_base_model", "training",
    "training.utils", "training.utils.train_utils", "training.utils.rag_e2e_dataloader_utils",
    "training.utils.retriever_only_dataloader_utils", "training.rag_e2e", "training.rag_e2e.train_rage2e",
    "training.retriever_only", "training.retriever_only.train_retriever_only",
    "eval", "eval.utils", "eval.eval_results", "eval.eval_retriever_only", "eval.eval_rag",
]


class _LazyAlias:
    """meta-path finder mapping dalm.X -> dalm_b200.X on first import"""

    @staticmethod
    def find_spec(name, path=None, target=None):
        if not name.startswith("dalm.") or name[5:] not in _ALIASES:
            return None
        real = importlib.import_module("dalm_b200." + name[5:])
        sys.modules[name] = real
        return real.__spec__


sys.meta_path.insert(0, _LazyAlias)
