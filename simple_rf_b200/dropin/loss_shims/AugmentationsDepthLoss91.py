from simple_rf_b200.loss_functions.AugmentationsDepthLoss91 import AugmentationsDepthLoss  # noqa: F401
