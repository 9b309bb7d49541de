from simple_rf_b200.models.SimpleTensoRF91 import SimpleTensoRF  # noqa: F401  (shim for models.ModelFactory02)
