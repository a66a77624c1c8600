"""`import RNA` resolves here when tests/golden/make_golden.py runs the reference CLI on the bpp parameter sets:
ViennaRNA is not installed, tests/fake_rna.py stands in for it (see its docstring)."""
from tests.fake_rna import fold_compound  # noqa: F401
