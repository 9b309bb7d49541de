"""Shim: lets the reference's unmodified DataPreprocessorFactory01 resolve `DataPreprocessor91` to simple_rf_b200's class."""
from simple_rf_b200.data_preprocessors.DataPreprocessor91 import DataPreprocessor  # noqa: F401
