"""gpirt_b200 — B200-native GP-IRT Gibbs sampler behind the reference's gpirtMCMC() interface.

Public surface mirrors the reference's NAMESPACE (gpirtMCMC, response_matrix, is.response_matrix,
as.response_matrix, data senate116); all sampling happens in the CUDA library (libgpirt_b200.so)."""
from .response_matrix import (ResponseMatrix, ResponseMessage, as_response_matrix, is_response_matrix,  # noqa: F401
                              response_matrix)
from .sampler import Sampler, gpirtMCMC  # noqa: F401
from .data import senate116  # noqa: F401
