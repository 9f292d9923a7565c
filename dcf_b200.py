"""Importable alias for the package directory, whose mandated name contains a '-' and therefore cannot
appear in an `import` statement:   import dcf_b200   ==   the package in
deep_continuous_fusion_for_multi-sensor_3d_object_detection_b200/ ."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "deep_continuous_fusion_for_multi-sensor_3d_object_detection_b200")
_spec = importlib.util.spec_from_file_location("dcf_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dcf_b200"] = _mod
_spec.loader.exec_module(_mod)
