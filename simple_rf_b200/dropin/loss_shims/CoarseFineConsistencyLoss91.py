from simple_rf_b200.loss_functions.CoarseFineConsistencyLoss91 import CoarseFineConsistencyLoss  # noqa: F401
