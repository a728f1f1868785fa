"""rspnet_b200 — B200-native (sm_100a) implementation of the RSPNet pretraining hot path.

Public surface mirrors the reference's plugin interface:
  rspnet_b200.models.get_model_class          (reference: models/__init__.py:16-75)
  rspnet_b200.moco.ModelFactory               (reference: moco/__init__.py:14-55)
  rspnet_b200.moco.builder_diffspeed_diffloss (reference: moco/builder_diffspeed_diffloss.py)
Compute goes through the C ABI in include/rspnet_b200.h; there is no CPU or library fallback.
"""
__version__ = "0.1.0"
