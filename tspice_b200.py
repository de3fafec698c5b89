"""Import alias: the package directory is `toy-spice_b200/` (not a valid identifier), so
`import tspice_b200` resolves to it."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("toy-spice_b200")
