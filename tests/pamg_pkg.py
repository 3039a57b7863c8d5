"""Imports the product package (its directory name is not a Python identifier)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
pamg = importlib.import_module("p-a_multigrids_b200")
