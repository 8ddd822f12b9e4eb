"""Import the UNMODIFIED upstream reference (read-only at /root/reference) with
its absent optional dependencies stubbed.

TEST INFRASTRUCTURE ONLY; used by oracle/make_golden.py and by the
container-only cross-check tests.  /root/reference does not exist on the GPU
box: nothing that runs there may call this.
"""
import os
import sys
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("SEASON_NERF_REFERENCE", "/root/reference")

_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.backends",
          "matplotlib.backends.backend_agg", "mpl_toolkits", "mpl_toolkits.axes_grid1", "hsluv", "gdal",
          "osgeo", "osgeo.gdal", "rpcm", "astropy", "astropy.coordinates", "astropy.time", "astropy.units",
          "robust_loss_pytorch", "maxflow", "sewar", "sewar.full_ref", "pyfftw", "pyfftw.interfaces",
          "pyfftw.interfaces.scipy_fftpack", "torch.utils.tensorboard"]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "T_NeRF_Full_2"))


def import_reference():
    """Returns a namespace with the reference symbols of the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    sys.dont_write_bytecode = True
    import numpy as _np
    if not hasattr(_np, "NaN"):          # the reference predates NumPy 2 (Quick_Run.py:38 uses np.NaN)
        _np.NaN = _np.nan
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for m in _STUBS:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    from types import SimpleNamespace
    import misc
    from T_NeRF_Full_2.T_NeRF_net_v2 import T_NeRF
    from T_NeRF_Full_2 import Eval_Tools_2
    from T_NeRF_Full_2.Quick_Run import Quick_Run_Net, encode_time
    from T_NeRF_Eval_Utils import mg_Img_Eval
    from all_NeRF.mg_unit_converter import world_angle_2_local_vec
    return SimpleNamespace(misc=misc, T_NeRF=T_NeRF, Eval_Tools_2=Eval_Tools_2,
                           All_in_One_Eval=Eval_Tools_2.All_in_One_Eval, get_PV=Eval_Tools_2.get_PV,
                           create_solor_rays_uniform=Eval_Tools_2.create_solor_rays_uniform,
                           Quick_Run_Net=Quick_Run_Net, encode_time=encode_time, mg_Img_Eval=mg_Img_Eval,
                           world_angle_2_local_vec=world_angle_2_local_vec)
