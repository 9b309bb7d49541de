"""Drop-in loss classes for the reference's unmodified LossComputer (src/loss_functions/LossComputer03.py:21-32):
module `<name>` -> class `<name>` minus its two-digit suffix."""
