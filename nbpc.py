"""Importable alias: `import nbpc` == the package in ./n-body_pointcloudevolution_b200 (whose
directory name, fixed by the project layout, is not a valid Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("n-body_pointcloudevolution_b200")
