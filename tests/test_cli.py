"""Render CLI (main_run_Season_NeRF.py, drop-in for the reference's file of the same name): flags, model-directory files
and - on the GPU - the rendered image against the CPU oracle."""
import json
import os
import sys

import numpy as np
import pytest
import torch as t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import main_run_Season_NeRF as cli  # noqa: E402


def test_cli_flags_match_the_reference():
    """main_run_Season_NeRF.py:10-44 of the reference: same flags, defaults and types."""
    a = cli.get_opts(["--Model_Location", "m", "--VA", "80", "0", "--SA", "45", "135", "--tf", "07/04"])
    assert a.Model_Location == "m" and a.VA == [80.0, 0.0] and a.SA == [45.0, 135.0] and a.tf == "07/04"
    assert tuple(a.Output_Size) == (256, 256, 96) and a.Save_Name is None
    assert a.ignore_progess is False and a.exact_shadow is False and a.Force_CPU is False
    b = cli.get_opts(["--Model_Location", "m", "--VA", "80", "0", "--SA", "45", "135", "--tf", "12/31", "--Output_Size", "8", "9",
                      "24", "--Save_Name", "x.png", "--ignore_progess", "--exact_shadow", "--Force_CPU"])
    assert tuple(b.Output_Size) == (8, 9, 24) and b.exact_shadow and b.Force_CPU and b.Save_Name == "x.png"
    assert cli.parse_time("01/01") == 0.0 and abs(cli.parse_time("07/04") - 184 / 365) < 1e-12     # :59-63


def test_cli_refuses_cpu(tmp_path):
    a = cli.get_opts(["--Model_Location", str(tmp_path), "--VA", "80", "0", "--SA", "45", "135", "--tf", "07/04", "--Force_CPU"])
    with pytest.raises(SystemExit):
        cli.render(a)


def _model_dir(tmp_path, params):
    from oracle import season_oracle as so
    json.dump({"fc_units": 512, "number_low_frequency_cases": 4, "n_samples": 96}, open(tmp_path / "opts.json", "w"))
    t.save({k: v.clone() for k, v in params.items()}, str(tmp_path / "Final_Model.nn"))
    np.save(str(tmp_path / "W2C_W2L_H.npy"), {"W2C": so.OMA_W2C, "W2L_H": so.oma_w2l_h()}, allow_pickle=True)
    return str(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [False, True])
def test_cli_render_matches_oracle(tmp_path, params0, exact):
    """checkpoint directory -> image, end to end (main_run_Season_NeRF.py:65-92) vs the CPU oracle (bf16 bar: 1e-2)."""
    from oracle import season_oracle as so
    loc = _model_dir(tmp_path, params0)
    argv = ["--Model_Location", loc, "--VA", "80", "0", "--SA", "45", "135", "--tf", "07/04", "--Output_Size", "6", "5", "96"]
    img, imgs, _ = cli.render(cli.get_opts(argv + (["--exact_shadow"] if exact else [])))
    size = (6, 5, 96)
    D = so.component_render_by_dir(params0, [80, 0], [45, 135], cli.parse_time("07/04"), size, so.OMA_W2C, so.oma_w2l_h(),
                                   include_exact_solar=exact)
    ref = so.get_imgs_from_img_dict(D, size)
    assert img.shape == (6, 5, 3)
    assert float(np.abs(img - ref["Season_Adj_Img"] * ref["Shadow_Adjust"]).max()) < 1e-2
    if exact:
        assert float(np.abs(imgs["Shadow_Mask_Exact"] - ref["Shadow_Mask_Exact"]).max()) < 2e-2
