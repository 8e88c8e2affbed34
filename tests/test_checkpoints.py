"""Checkpoint bridge: flat buffers <-> flax msgpack `checkpoint_<step>` <-> torch `*.ckpt`.

Pinned by tests/golden/ckpt_bridge.npz: make_golden.py wrote a checkpoint with this module and let the REFERENCE's
own restore_model_state_from_jaxnerf (octree/nerf/models.py:66-113) load it into the reference torch model; the
fixture holds the sha256 of those bytes, the reference model's state_dict keys / shapes / sums and its
eval_points_raw outputs."""
import hashlib
import os
import types

import numpy as np
import torch

from oracle import nerf_sh_oracle as O
from plenoctree_b200.nerf import checkpoints as C


def _inputs(golden_dir):
    z = np.load(os.path.join(golden_dir, "ckpt_bridge.npz"))
    sh_deg = int(z["sh_deg"])
    s0, s1, s2 = [int(s) for s in z["seeds"]]
    flat = np.concatenate([O.init_flat_params(sh_deg, s0, bias_scale=0.05), O.init_flat_params(sh_deg, s1, bias_scale=0.05)])
    rs = np.random.RandomState(s2)
    m = rs.normal(size=flat.shape).astype(np.float32) * 1e-3
    v = (rs.uniform(size=flat.shape).astype(np.float32) * 1e-6)
    return z, sh_deg, flat, m, v, int(z["step"])


def test_bytes_are_the_ones_the_reference_loader_consumed(golden_dir):
    z, sh_deg, flat, m, v, step = _inputs(golden_dir)
    blob = C.msgpack_serialize(C.train_state_dict(flat, m, v, step, sh_deg))
    assert len(blob) == int(z["nbytes"])
    assert hashlib.sha256(blob).hexdigest() == str(z["sha256"])
    sd = C.msgpack_restore(blob)
    assert set(sd["optimizer"].keys()) == {"target", "state"}
    assert sd["optimizer"]["target"]["params"]["MLP_1"]["Dense_5"]["kernel"].shape == (319, 256)
    assert sd["optimizer"]["target"]["params"]["MLP_0"]["Dense_9"]["kernel"].shape == (256, 48)
    p2, m2, v2, step2 = C.state_dict_to_flat(sd, sh_deg)
    assert step2 == step and (p2 == flat).all() and (m2 == m).all() and (v2 == v).all()


def test_torch_state_dict_matches_reference_model(golden_dir):
    z, sh_deg, flat, _, _, _ = _inputs(golden_dir)
    sd = C.flat_to_torch_state_dict(flat, sh_deg)
    keys = [str(k) for k in z["keys"]]
    assert sorted(sd.keys()) == keys
    for k, shp, s, a in zip(keys, z["shapes"], z["sums"], z["abs_sums"]):
        assert ";".join(map(str, sd[k].shape)) == str(shp), k
        assert abs(float(sd[k].astype(np.float64).sum()) - float(s)) <= 1e-9 * max(1.0, float(a)), k
        assert abs(float(np.abs(sd[k].astype(np.float64)).sum()) - float(a)) <= 1e-9 * max(1.0, float(a)), k
    assert (C.torch_state_dict_to_flat(sd, sh_deg) == flat).all()


def test_reference_model_outputs_match_oracle_on_restored_params(golden_dir):
    z, sh_deg, flat, m, v, step = _inputs(golden_dir)
    p2, _, _, _ = C.state_dict_to_flat(C.msgpack_restore(C.msgpack_serialize(C.train_state_dict(flat, m, v, step, sh_deg))), sh_deg)
    P = C.param_count(sh_deg)
    pts = torch.from_numpy(z["points"])
    with torch.no_grad():
        rgb_c, sig_c = O.eval_points_raw(O.unflatten(p2[:P], sh_deg), pts)
        rgb_f, sig_f = O.eval_points_raw(O.unflatten(p2[P:], sh_deg), pts)
    for got, want in ((rgb_c, z["raw_rgb_coarse"]), (sig_c, z["raw_sigma_coarse"]), (rgb_f, z["raw_rgb_fine"]),
                      (sig_f, z["raw_sigma_fine"])):
        assert np.abs(got.numpy() - want).max() <= 2e-5 * max(1.0, np.abs(want).max())


class _FakeModel:
    def __init__(self, sh_deg, flat):
        self.sh_deg = sh_deg
        self.params = torch.from_numpy(flat.copy())

    def set_params(self, flat):
        self.params = torch.as_tensor(flat, dtype=torch.float32).clone()


def test_save_restore_files_latest_and_keep(tmp_path, golden_dir):
    _, sh_deg, flat, m, v, _ = _inputs(golden_dir)
    model = _FakeModel(sh_deg, flat)
    state = types.SimpleNamespace(m=torch.from_numpy(m.copy()), v=torch.from_numpy(v.copy()), step=0)
    d = str(tmp_path / "train")
    for step in (5, 9, 10, 100):
        state.step = step
        model.params[0] = float(step)
        C.save_checkpoint(d, model, state, keep=3)
    names = sorted(os.listdir(d))
    assert names == ["checkpoint_10", "checkpoint_100", "checkpoint_9"]      # keep=3 dropped checkpoint_5
    assert os.path.basename(C.latest_checkpoint(d)) == "checkpoint_100"      # natural, not lexicographic, order
    model2 = _FakeModel(sh_deg, np.zeros_like(flat))
    state2 = types.SimpleNamespace(m=torch.zeros(flat.size), v=torch.zeros(flat.size), step=0)
    assert C.restore_checkpoint(d, model2, state2) == 100
    assert float(model2.params[0]) == 100.0 and (model2.params[1:].numpy() == flat[1:]).all()
    assert (state2.m.numpy() == m).all() and (state2.v.numpy() == v).all() and state2.step == 100
    assert C.restore_checkpoint(str(tmp_path / "empty"), model2, state2) is None
    # parameters-only load, like octree.extraction --is_jaxnerf_ckpt
    model3 = _FakeModel(sh_deg, np.zeros_like(flat))
    assert C.restore_model_state_from_jaxnerf(d, model3) and float(model3.params[0]) == 100.0


def test_torch_ckpt_round_trip(tmp_path, golden_dir):
    _, sh_deg, flat, _, _, _ = _inputs(golden_dir)
    d = str(tmp_path)
    C.save_torch_ckpt(os.path.join(d, "000100.ckpt"), _FakeModel(sh_deg, flat))
    ck = torch.load(os.path.join(d, "000100.ckpt"), map_location="cpu")
    assert set(ck.keys()) == {"model"} and ck["model"]["MLP_1.rgb_layer.weight"].shape == (48, 256)
    model = _FakeModel(sh_deg, np.zeros_like(flat))
    assert C.restore_model_state(d, model).endswith("000100.ckpt")
    assert (model.params.numpy() == flat).all()


def test_wrong_sh_deg_is_rejected(golden_dir):
    _, sh_deg, flat, m, v, step = _inputs(golden_dir)
    sd = C.train_state_dict(flat, m, v, step, sh_deg)
    try:
        C.state_dict_to_flat(sd, 4)
    except ValueError as e:
        assert "Dense_9" in str(e)
    else:
        raise AssertionError("SH16 checkpoint loaded as SH25")
