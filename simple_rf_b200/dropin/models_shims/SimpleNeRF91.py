from simple_rf_b200.models.SimpleNeRF91 import SimpleNeRF  # noqa: F401  (shim for models.ModelFactory02)
