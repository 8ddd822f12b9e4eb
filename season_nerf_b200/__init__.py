"""season_nerf_b200: B200-native (sm_100a) implementation of the Season-NeRF render / train hot path.

Public surface mirrors the reference's Python API for this path:
  T_NeRF, G_NeRF_Net_Classic, SineLayer, PE_Encode          <- T_NeRF_Full_2/T_NeRF_net_v2.py, G_NeRF.py, misc.py
  All_in_One_Eval, get_PV, create_solor_rays_uniform          <- T_NeRF_Full_2/Eval_Tools_2.py
  sample_pt_coarse, zero_invalid_pts                          <- misc.py
  component_render_by_dir/_by_P, _internal_render,
  get_imgs_from_Img_Dict, get_imgs_from_Img_Dict_t_step       <- T_NeRF_Eval_Utils/mg_Img_Eval.py
  Quick_Run_Net, encode_time                                  <- T_NeRF_Full_2/Quick_Run.py
  T_NeRF_Net_Tool, Net_tool                                   <- T_NeRF_Full_2/Net_Tool_2.py, mg_run_NeRF.py
  world_angle_2_local_vec                                     <- all_NeRF/mg_unit_converter.py
`season_nerf_b200.compat.install()` registers these under the reference's module names.
"""
from .adaptive_loss import AdaptiveLossFunction
from .align import Grad_Descent_Seasonal_Align_v3
from .data import RayTable, data_to_dict
from .engine import (All_in_One_Eval, create_solor_rays_uniform, get_PV, sample_pt_coarse, sample_ts,
                     zero_invalid_pts)
from .geometry import LLA_get_vec, encode_time, ray_table_from_P, world_angle_2_local_vec
from .net_tool import ColorTable, Net_tool, T_NeRF_Net_Tool
from .network import G_NeRF_Net_Classic, PE_Encode, SineLayer, T_NeRF
from .quick_run import Quick_Run_Net
from .render import (DeviceImgDict, _internal_render, component_render_by_dir, component_render_by_P,
                     get_imgs_from_Img_Dict, get_imgs_from_Img_Dict_t_step, render_image_sharded, render_shard)
from .train import TrainStep
from .shadow_eval import Test_Shadow_Points, eval_shadow_data, shadow_anaylysis
from .volume import confidence_range, eval_HM, gen_results, height_map

__version__ = "0.1.0"
