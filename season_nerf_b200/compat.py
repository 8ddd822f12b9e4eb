"""Drop-in installer: makes the reference's own import statements resolve to this package for the hot path.

    import season_nerf_b200.compat as c; c.install()          # before importing the reference's modules
    from T_NeRF_Full_2.T_NeRF_net_v2 import T_NeRF            # -> season_nerf_b200.network.T_NeRF
    from T_NeRF_Full_2.Eval_Tools_2 import All_in_One_Eval    # -> season_nerf_b200.engine.All_in_One_Eval
    from T_NeRF_Eval_Utils.mg_Img_Eval import component_render_by_dir, get_imgs_from_Img_Dict

Modules that exist only for this path are replaced wholesale (T_NeRF_net_v2, G_NeRF, Eval_Tools_2, Quick_Run);
`misc` and `mg_Img_Eval` hold unrelated helpers too, so when the reference tree is importable their hot-path symbols
are patched in place and everything else is left alone; otherwise minimal stand-in modules are registered.
"""
import importlib
import sys
import types

_HOT = {
    "T_NeRF_Full_2.T_NeRF_net_v2": ("network", ["T_NeRF", "G_NeRF_Net_Classic", "SineLayer", "PE_Encode"]),
    "T_NeRF_Full_2.G_NeRF": ("network", ["G_NeRF_Net_Classic", "SineLayer", "PE_Encode"]),
    "T_NeRF_Full_2.Eval_Tools_2": ("engine", ["All_in_One_Eval", "get_PV", "create_solor_rays_uniform", "sample_pt_coarse"]),
    "T_NeRF_Full_2.Quick_Run": ("quick_run", ["Quick_Run_Net", "encode_time", "All_in_One_Eval", "create_solor_rays_uniform"]),
}
# the training driver: `from T_NeRF_Full_2.Net_Tool_2 import T_NeRF_Net_Tool` (main.py:12) -> season_nerf_b200.net_tool
_NET_TOOL = {"T_NeRF_Full_2.Net_Tool_2": ("net_tool", ["T_NeRF_Net_Tool", "Net_tool", "T_NeRF", "All_in_One_Eval", "AdaptiveLossFunction",
                                                        "get_output_loc_lin_first"])}
_PATCH = {
    "misc": ("engine", ["sample_pt_coarse", "zero_invalid_pts"], "network", ["SineLayer", "PE_Encode"]),
    "T_NeRF_Eval_Utils.mg_Img_Eval": ("render", ["_internal_render", "component_render_by_dir", "component_render_by_P",
                                                 "get_imgs_from_Img_Dict", "get_imgs_from_Img_Dict_t_step"]),
    "all_NeRF.mg_unit_converter": ("geometry", ["world_angle_2_local_vec", "LLA_get_vec"]),
    "mg_run_NeRF": ("net_tool", ["Net_tool"]),
    "T_NeRF_Eval_Utils.mg_Shadow_Eval": ("shadow_eval", ["eval_shadow_data", "Test_Shadow_Points", "shadow_anaylysis"]),
    "T_NeRF_Eval_Utils.Eval_funcs": ("volume", ["gen_results"]),
}


def _ours(modname):
    return importlib.import_module("season_nerf_b200." + modname)


def install(patch_existing=True, net_tool="ours"):
    """Returns the list of module names that now resolve to season_nerf_b200.
    net_tool="ours" (default): `T_NeRF_Full_2.Net_Tool_2` resolves to season_nerf_b200.net_tool, whose T_NeRF_Net_Tool runs
    every section's step - optimiser updates included - from one captured CUDA graph.
    net_tool="reference": the reference's own Net_Tool_2.py is left to load; its imports (`T_NeRF`, `All_in_One_Eval`,
    `mg_run_NeRF.Net_tool`, `AdaptiveLossFunction`) resolve here, so the UNMODIFIED reference class drives this package:
    its reset_eval builds the eval tool / optimisers, our Net_tool.train_step adopts them (graph for forward + backward)."""
    if net_tool not in ("ours", "reference"):
        raise ValueError("net_tool must be 'ours' or 'reference'")
    done = []
    try:
        importlib.import_module("robust_loss_pytorch")
    except Exception:
        m = types.ModuleType("robust_loss_pytorch")
        m.__doc__ = "season_nerf_b200.adaptive_loss standing in for the absent robust_loss_pytorch package"
        m.AdaptiveLossFunction = _ours("adaptive_loss").AdaptiveLossFunction
        m.__dict__["__season_nerf_b200__"] = True
        sys.modules["robust_loss_pytorch"] = m
        done.append("robust_loss_pytorch")
    hot = dict(_HOT)
    if net_tool == "ours":
        hot.update(_NET_TOOL)
    for name, (src, symbols) in hot.items():
        m = types.ModuleType(name)
        m.__doc__ = "season_nerf_b200 drop-in for the reference module " + name
        o = _ours(src)
        for s in symbols:
            m.__dict__[s] = getattr(o, s)
        m.__dict__["__season_nerf_b200__"] = True
        sys.modules[name] = m
        parent = sys.modules.get(name.rsplit(".", 1)[0])
        if parent is not None:
            setattr(parent, name.rsplit(".", 1)[1], m)
        done.append(name)
    for name, spec in _PATCH.items():
        target = sys.modules.get(name)
        if target is None and patch_existing:
            try:
                target = importlib.import_module(name)
            except Exception:
                target = None
        if target is None:
            target = types.ModuleType(name)
            sys.modules[name] = target
        for i in range(0, len(spec), 2):
            o = _ours(spec[i])
            for s in spec[i + 1]:
                setattr(target, s, getattr(o, s))
        done.append(name)
    return done
