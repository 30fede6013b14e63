"""Import shim: the package directory is named `gpu-accel-ofdm-ls-mrc_b200` (hyphens, as
the task names it), which Python cannot import by name.  `import ofdm_b200` loads it
from its path and registers it as `gpu_accel_ofdm_ls_mrc_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu-accel-ofdm-ls-mrc_b200")
_NAME = "gpu_accel_ofdm_ls_mrc_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
CONFIGS = pkg.CONFIGS
RxConfig = pkg.RxConfig
LsMrcReceiver = pkg.LsMrcReceiver
LsmrcError = pkg.LsmrcError
load_library = pkg.load_library
synth = pkg.synth
build = pkg.build
sharding = pkg.sharding
ABI = pkg.ABI
