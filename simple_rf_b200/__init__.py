"""simple_rf_b200 — B200 (sm_100a) kernels for the per-ray rendering hot path of Simple-RF.

`ops` wraps the C ABI (include/simple_rf_b200.h); `models` holds the drop-in model classes that the
reference's unmodified ModelFactory02 / Trainer10 / Tester07 can drive.  There is no CPU fallback.
"""
__version__ = '0.1.0'
