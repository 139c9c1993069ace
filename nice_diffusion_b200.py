"""Import shim: the product package lives in the directory ``nice-diffusion_b200/`` (a name Python cannot
import directly). Importing ``nice_diffusion_b200`` loads that directory as a regular package under this name:

    from nice_diffusion_b200.model import DiffusionModel
    from nice_diffusion_b200.diffusion import Diffusion
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'nice-diffusion_b200')
_spec = importlib.util.spec_from_file_location(
    'nice_diffusion_b200', os.path.join(_pkg_dir, '__init__.py'), submodule_search_locations=[_pkg_dir])
_module = importlib.util.module_from_spec(_spec)
sys.modules['nice_diffusion_b200'] = _module
_spec.loader.exec_module(_module)
