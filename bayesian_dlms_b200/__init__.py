"""bayesian_dlms_b200 -- B200-native Kalman hot path of jonnylaw/bayesian_dlms.

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI of ``include/bdlm.h``),
the ctypes binding (``_capi``), the batched host shim (``batch``), the host-side model
assembly (``dlm``) and a mirror of the reference's operator API (``reference_api``).
The compute path is libbdlm.so only; importing this package never falls back to a CPU
implementation, and nothing here imports ``oracle/``.
"""
from .dlm import (Data, Dlm, DlmParameters, autoregressive, polynomial, regression,  # noqa: F401
                  seasonal)
from .batch import (Engine, Model, SERIES_MAJOR, TIME_MAJOR, build_batch_model,  # noqa: F401
                    default_engine, model_to_device)
from .reference_api import (ConjugateFilter, FilterAr, FilterOu, GibbsSampling,  # noqa: F401
                            InverseGamma, InverseWishart, KalmanFilter, KfState, SamplingState,
                            Smoothing, SmoothingState, SvParameters, SvdFilter, SvdSampler,
                            SvdState)

__all__ = [
    "Data", "Dlm", "DlmParameters", "polynomial", "regression", "autoregressive", "seasonal",
    "Engine", "Model", "TIME_MAJOR", "SERIES_MAJOR", "default_engine", "build_batch_model",
    "model_to_device",
    "KalmanFilter", "Smoothing", "SvdFilter", "SvdSampler", "GibbsSampling",
    "KfState", "SmoothingState", "SamplingState", "SvdState",
    "FilterAr", "FilterOu", "SvParameters", "ConjugateFilter", "InverseGamma", "InverseWishart",
]
