"""exemplars_vc_b200 -- B200-native activation estimation for exemplar-based voice conversion.

One hot path of entn-at/exemplars_vc, rebuilt for sm_100a: sparse non-negative activations H of a
spectrogram X over a fixed exemplar dictionary A by multiplicative updates
``H <- H * A^T(X / AH) / (A^T 1 + lambda)`` and the conversion product ``Y = B H``
(reference: 04_align_n_nmf.py:194-215, 336-393; nmf_tool/nmf.py; 05_conversion.py).

Public surface (mirrors the reference's entry points):
  nmf.non_negative_factorization      -- the operator the reference calls (scikit-learn's signature)
  align_n_nmf._factorize / factorize / convert          (04_align_n_nmf.py)
  align_n_nmf_pytorch._factorize / factorize / convert  (04_align_n_nmf_pytorch.py)
  conversion                                            (05_conversion.py)
  nmf_tool.nmf.NMF                                      (nmf_tool/nmf.py)
  ExemplarDictionary                  -- the device-resident dictionary behind all of them
  sharding                            -- utterance / exemplar sharding over torch.distributed (NCCL)

The arithmetic lives in libevc_b200.so (hand-written CUDA, C ABI in include/evc.h).  There is no
CPU implementation in this package: without the built library or without a GPU the calls raise.
"""
from .dictionary import Activation, ExemplarDictionary  # noqa: F401
from .nmf import ConvergenceWarning, non_negative_factorization  # noqa: F401

__all__ = ["ExemplarDictionary", "Activation", "non_negative_factorization", "ConvergenceWarning"]
__version__ = "0.1.0"
