"""Host-side engine: parameter layout in HBM and the per-layer launch sequence of the encoder / decoder forward and
backward, expressed over the C-ABI kernels (dalm_b200.ops). No arithmetic happens in PyTorch here."""
