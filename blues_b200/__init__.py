"""blues_b200 — B200-native NCMC engine behind the BLUES Python API."""
__version__ = '0.1.0'
