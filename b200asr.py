"""Alias so that `import b200asr` resolves to the package in `end-to-end-asr-pytorch_b200/`."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("end-to-end-asr-pytorch_b200")
