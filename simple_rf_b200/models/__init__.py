"""Drop-in model classes (same public surface as the reference's src/models/SimpleNeRF17.py and
SimpleTensoRF09.py), evaluated by the sm_100a kernels."""
