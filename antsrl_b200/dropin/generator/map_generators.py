"""generator/map_generators.py of the reference (lines 9-46)."""
from antsrl_b200.generator import CirclesGenerator, PerlinGenerator   # noqa: F401
