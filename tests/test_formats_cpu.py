"""§8(f)4: the reference's on-disk formats (checkpoint list, PBR-extended PLY) round-trip, and the PLY written here is
readable by a reader that follows GaussianModel.load_ply's access pattern (scene/gaussian_model.py:568-666)."""
import numpy as np
import torch

from relightable3dgaussian_b200 import formats


def _model(P=37, pbr=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    m = dict(xyz=r(P, 3), normal=r(P, 3), f_dc=r(P, 1, 3), f_rest=r(P, 15, 3), opacity=r(P, 1), scaling=r(P, 3), rotation=r(P, 4))
    if pbr:
        m.update(base_color=r(P, 3), roughness=r(P, 1), incidents_dc=r(P, 1, 3), incidents_rest=r(P, 15, 3),
                 visibility_dc=r(P, 1, 1), visibility_rest=r(P, 15, 1))
    return m


def test_attribute_names_follow_the_reference():
    names = formats.attribute_names(_model())
    # construct_list_of_attributes (gaussian_model.py:507-533): order and naming
    assert names[:6] == ["x", "y", "z", "nx", "ny", "nz"] and names[6:9] == ["f_dc_0", "f_dc_1", "f_dc_2"]
    assert names[9] == "f_rest_0" and names[9 + 44] == "f_rest_44" and names[54] == "opacity"
    assert names[55:58] == ["scale_0", "scale_1", "scale_2"] and names[58:62] == ["rot_0", "rot_1", "rot_2", "rot_3"]
    assert names[62:65] == ["base_color_0", "base_color_1", "base_color_2"] and names[65] == "roughness"
    assert names[66] == "incidents_dc_0" and names[69] == "incidents_rest_0" and names[114] == "visibility_dc_0"
    assert names[-1] == "visibility_rest_14" and len(names) == 62 + 3 + 1 + 48 + 16
    assert len(formats.attribute_names(_model(pbr=False))) == 62


def test_ply_round_trip_and_layout(tmp_path):
    for pbr in (True, False):
        m = _model(pbr=pbr, seed=1)
        path = str(tmp_path / f"pc_{pbr}" / "point_cloud.ply")
        formats.save_ply(path, m)
        back = formats.load_ply(path)
        assert set(back) == set(m)
        for k in m:
            assert back[k].shape == m[k].shape and torch.equal(back[k], m[k]), k
        # channel-major flattening (save_ply: transpose(1, 2).flatten(1)): f_rest_0..14 are channel 0's coefficients
        raw = open(path, "rb").read()
        hdr_end = raw.index(b"end_header\n") + len(b"end_header\n")
        assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 37\nproperty float x\n")
        data = np.frombuffer(raw[hdr_end:], "<f4").reshape(37, -1)
        names = formats.attribute_names(m)
        assert data.shape[1] == len(names)
        assert np.array_equal(data[:, names.index("f_rest_0")], m["f_rest"][:, 0, 0].numpy())
        assert np.array_equal(data[:, names.index("f_rest_15")], m["f_rest"][:, 0, 1].numpy())
        assert np.array_equal(data[:, names.index("f_dc_2")], m["f_dc"][:, 0, 2].numpy())
        # the reference's reader addresses properties BY NAME (plydata.elements[0]["f_dc_1"], sorted f_rest_*): emulate it
        col = {n: data[:, i] for i, n in enumerate(names)}
        rest = np.stack([col[f"f_rest_{i}"] for i in range(45)], 1).reshape(37, 3, 15)          # gaussian_model.py:590
        assert np.array_equal(np.transpose(rest, (0, 2, 1)), m["f_rest"].numpy())
    # ascii PLYs (other writers) load too
    p = tmp_path / "a.ply"
    m = _model(P=3, pbr=False)
    names = formats.attribute_names(m)
    rows = np.concatenate([m[k].transpose(1, 2).reshape(3, -1).numpy() if k in formats._CHANNEL_MAJOR else m[k].reshape(3, -1).numpy()
                           for k in formats.GEOMETRY], 1)
    p.write_text("ply\nformat ascii 1.0\ncomment test\nelement vertex 3\n" + "".join(f"property float {n}\n" for n in names) +
                 "end_header\n" + "\n".join(" ".join(repr(float(v)) for v in r) for r in rows) + "\n")
    back = formats.load_ply(str(p))
    assert torch.allclose(back["f_rest"], m["f_rest"]) and torch.allclose(back["rotation"], m["rotation"])


def test_checkpoint_list_layout_round_trip(tmp_path):
    m = _model(seed=2)
    P = m["xyz"].shape[0]
    stats = dict(max_radii2D=torch.zeros(P), weights_accum=torch.rand(P, 1), xyz_gradient_accum=torch.rand(P, 1),
                 normal_gradient_accum=torch.rand(P, 1), denom=torch.ones(P, 1))
    params = [torch.nn.Parameter(v.clone()) for v in m.values()]
    opt = torch.optim.Adam([{"params": [p], "lr": 1e-3, "name": k} for p, k in zip(params, m)], lr=0.0, eps=1e-15)
    cap = formats.capture(m, stats, opt.state_dict(), active_sh_degree=3, spatial_lr_scale=2.5)
    assert len(cap) == 21 and cap[0] == 3 and cap[1] is m["xyz"] and cap[7] is m["opacity"] and cap[14] == 2.5    # capture():114-142
    assert cap[5] is m["scaling"] and cap[6] is m["rotation"] and cap[15] is m["base_color"] and cap[20] is m["visibility_rest"]
    path = str(tmp_path / "chkpnt30000.pth")
    formats.save_checkpoint(path, cap, 30000)
    cap2, it = formats.load_checkpoint(path)
    model, st, opt_state, deg, scale = formats.restore(cap2)
    assert it == 30000 and deg == 3 and scale == 2.5 and set(model) == set(m) and set(st) == set(stats)
    for k in m:
        assert torch.equal(model[k], m[k])
    assert [g["name"] for g in opt_state["param_groups"]] == list(m)
    assert len(formats.capture({k: m[k] for k in formats.GEOMETRY}, stats, {}, 3, 1.0)) == 15


def test_bake_persistence(tmp_path):
    vis = (torch.rand(11, 8, 1) > 0.4).float() * (0.9 + 0.1 * torch.rand(11, 8, 1))
    dirs = torch.nn.functional.normalize(torch.randn(11, 8, 3), dim=-1)
    p = str(tmp_path / "bake.pth")
    formats.save_bake(p, vis, dirs, torch.full((11, 8, 1), 2 * np.pi))
    v, d, a = formats.load_bake(p)
    assert torch.equal(v, vis) and torch.equal(d, dirs) and torch.allclose(a, torch.full_like(a, 2 * np.pi))
