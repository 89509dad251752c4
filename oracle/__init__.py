"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / torch-fp32/fp64 on CPU) of the reference's algorithm for the RAG-e2e / retriever-only
training step. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import anything from here, and only as the checker / the reported CPU baseline — never as a product path.

Parity pinning: the reference's own tests for this path are `assert True` stubs (reference
tests/training/rag_e2e/test_base_model.py:1-2, test_train_rage2e.py:1-2), so there are no upstream golden vectors.
The restatements in `oracle/losses.py`, `oracle/pooling.py`, `oracle/preprocess.py` are instead pinned against OUTPUTS OF
THE REFERENCE ITSELF, executed in the build container by `oracle/make_golden.py` (which imports the unmodified reference
from /root/reference with `peft`/`accelerate` stubbed) and committed under `tests/golden/`.
Arithmetic that lives in third-party packages absent from the reference tree:
  * transformers (BertModel / LlamaForCausalLM forward) — installed here (5.5.0) and used directly as the oracle;
  * peft (LoRA, unpinned in the reference's pyproject.toml:16-33, not installed) — restated in `oracle/models.py`
    from the published LoRA definition with the reference's hyper-parameters: "parity unpinned" for that piece;
  * accelerate (DDP wrap / dataloader sharding, unpinned, not installed) — restated in tests against
    torch DistributedDataParallel semantics on gloo: "parity unpinned".
"""
