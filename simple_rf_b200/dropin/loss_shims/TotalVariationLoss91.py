from simple_rf_b200.loss_functions.TotalVariationLoss91 import TotalVariationLoss  # noqa: F401
