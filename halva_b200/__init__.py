"""halva_b200: B200-native (sm_100a) implementation of HALVA's DPA-loss hot path.

Public API
  dpa_loss(...)                  fused lm_head + log-softmax + phrase-masked gather + KL, forward and backward
  fused_compute_loss(...)        drop-in for HalvaTrainer.compute_loss (halva_b200.trainer); patch_halva_trainer(cls)
  halva_b200.dist                data-parallel / vocab-parallel helpers (average_gradients, vocab_shard_bounds, init_library_comm)
  halva_b200.host.HostStager     host-buffer entry: copies only the hidden rows that can contribute
"""
from .dpa_loss import DPALossFn, debug_gemm, dpa_loss, read_status  # noqa: F401
from ._lib import DPAError  # noqa: F401
from .trainer import fused_compute_loss, patch_halva_trainer  # noqa: F401

__all__ = ["dpa_loss", "DPALossFn", "DPAError", "read_status", "debug_gemm", "fused_compute_loss", "patch_halva_trainer"]
