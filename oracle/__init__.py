"""CPU oracle for the Kalman hot path (TEST INFRASTRUCTURE -- never imported by the product).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``bayesian_dlms_b200`` must not.
"""
from .oracle import (  # noqa: F401
    build, lib, kf_filter, rts_smooth, backward_sample, ffbs, loglik, svd_filter,
    svd_backward_sample, svd_ffbs, svd_filter_tv, svd_ffbs_tv, sqrt_svd, gibbs_stats, eigsym, svd, solve,
    batch_filter_smooth, batch_ffbs, gibbs_invgamma, inverse_wishart, ar_filter,
    ar_backward_sample, conjugate_filter,
)
