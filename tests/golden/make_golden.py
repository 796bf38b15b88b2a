"""Regenerate tests/golden/ from the reference checkout (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

* ``first_order_dlm*.csv`` are verbatim copies of the reference's committed example
  data and of the outputs its own apps wrote (``FilterDlm`` / ``SmoothDlm``,
  examples/src/main/scala/dlm/FirstOrderDlm.scala:52-77,237-255): input series, filtered
  (time, m, C, f, Q) and smoothed (time, s, S) at full fp64 print precision.
* ``kat.json`` transcribes the known-answer tables of the reference's unit tests
  (core/src/test/scala/KalmanFilter.scala:79-94,127-188, Smoothing.scala:12-28,
  SvdFilter.scala:105-122) -- numbers typed from those files, tolerance as stated there.
Nothing under /root/reference is read at test time.
"""
import json
import os
import shutil
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))
for name in ("first_order_dlm.csv", "first_order_dlm_filtered.csv",
             "first_order_dlm_smoothed.csv",
             # "next" rows: FilterAr.filterUnivariate output of FilterArDlm
             # (examples/src/main/scala/dlm/ar.scala:47-60) and ConjugateFilter output of
             # ConjFilter (FirstOrderDlm.scala:144-172)
             "ar_dlm.csv", "ar_dlm_filtered.csv", "first_order_dlm_conjugate_filtered.csv"):
    shutil.copyfile(os.path.join(ref, "examples", "data", name), os.path.join(here, name))

N = None
kat = {
    "kalman_filter_test": {
        "source": "core/src/test/scala/KalmanFilter.scala:79-94,127-188",
        "model": "polynomial(1) |*| polynomial(1)",
        "v_diag": [3.0, 3.0], "w_diag": [1.0, 1.0], "m0": [0.0, 0.0], "c0_diag": [1.0, 1.0],
        "times": [1.0, 2.0, 3.0, 4.0, 5.0, 7.0],
        "obs": [[4.5, 4.5], [3.0, 3.0], [6.3, 6.3], [N, N], [10.1, N], [15.2, 15.2]],
        "tol": 1e-4,
        # index = position in `times`; diagonal entries of the 2x2 matrices
        "expected": {
            "1": {"a": [1.8, 1.8], "R": [2.2, 2.2], "f": [1.8, 1.8], "Q": [5.2, 5.2],
                  "m": [2.307692, 2.307692], "C": [1.269231, 1.269231]},
            "2": {"a": [2.307692, 2.307692], "R": [2.269231, 2.269231],
                  "f": [2.307692, 2.307692], "Q": [5.269231, 5.269231],
                  "m": [4.027007, 4.027007], "C": [1.291971, 1.291971]},
            "3": {"a": [4.027007, 4.027007], "R": [2.291971, 2.291971],
                  "f": [4.027007, 4.027007], "Q": [5.291971, 5.291971],
                  "m": [4.027007, 4.027007], "C": [2.291971, 2.291971]},
            "4": {"a": [4.027007, 4.027007], "R": [3.291971, 3.291971],
                  "f": [4.027007, 4.027007], "Q": [6.291971, 6.291971],
                  "m": [7.204408, 4.027007], "C": [1.569606, 3.291971]},
            "5": {"a": [7.204408, 4.027007], "R": [3.569606, 5.291971],
                  "f": [7.204408, 4.027007], "Q": [6.569606, 8.291971],
                  # commented-out expectations at :187-188 (first component)
                  "m0_commented": 11.54883, "C00_commented": 1.630055},
        },
    },
    "smoothing_test": {
        "source": "core/src/test/scala/Smoothing.scala:12-28,42-75",
        "model": "polynomial(1)", "v": 3.0, "w": 1.0, "m0": 0.0, "c0": 1.0,
        "times": [1.0, 2.0, 3.0, 4.0, 5.0, 7.0],
        "obs": [4.5, 3.0, 6.3, N, 10.1, 15.2],
        "tol": 1e-4,
    },
    "svd_filter_test": {
        "source": "core/src/test/scala/SvdFilter.scala:105-157",
        "note": "same model/data as kalman_filter_test; SVD filter means (tol 1e-2) and "
                "uc diag(dc^2) uc^T (tol 1e-2) equal the Kalman filter's",
        "tol": 1e-2,
    },
}
with open(os.path.join(here, "kat.json"), "w") as fh:
    json.dump(kat, fh, indent=1)
print("golden fixtures written to", here)
