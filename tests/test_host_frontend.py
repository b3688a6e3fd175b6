"""Host front-end: config grammar (PPMGenerator.hpp:174-618), error behaviour, PPM
writer, camera — against what the reference executable itself printed
(tests/golden/cli_errors.json, produced by tools/gen_golden.py)."""
import json

import numpy as np
import pytest

import oracle_bindings as ob
from conftest import GOLD
from whittedstyle_raytracer_b200 import Scene, SceneError, fixtures, read_ppm_p3, write_ppm_p3

CLI = json.loads((GOLD / "cli_errors.json").read_text())
HEADER = "imsize 4 4\neye 0 0 0\nviewdir 0 0 -1\nupdir 0 1 0\nhfov 60\nbkgcolor 0 0 0 1\n"


def _expected_text(err_line):
    # the reference prints "ERROR: " + what() for exceptions, and a few lines directly as "ERROR:: ..."
    return err_line if err_line.startswith("ERROR::") else err_line[len("ERROR: "):]


@pytest.mark.parametrize("case", [k for k, v in CLI.items() if v["returncode"] == 255 and k != "face_index_oob"])
def test_error_text_matches_reference(case, workdir):
    c = CLI[case]
    with pytest.raises(SceneError) as e:
        Scene(text=c["config"], asset_dir=workdir)
    assert str(e.value).splitlines()[0] == _expected_text(c["error_line"])


def test_out_of_range_face_index_is_an_error(workdir):
    # the reference builds its message with `int + const char*` pointer arithmetic (UB); only the failure is pinned
    with pytest.raises(SceneError):
        Scene(text=CLI["face_index_oob"]["config"], asset_dir=workdir)
    assert CLI["face_index_oob"]["returncode"] == 255


@pytest.mark.parametrize("case", ["ok_minimal", "last_keyword_no_newline"])
def test_ppm_bytes_match_reference(case, workdir, tmp_path):
    """Reference executable's .ppm vs oracle render + our P3 writer: byte-identical."""
    c = CLI[case]
    assert c["returncode"] == 0
    scene = Scene(text=c["config"], asset_dir=workdir)
    assert scene.desc.shadow_type == 0          # a trailing `shadow` keyword without newline is dropped, :179-183
    img, _ = ob.OracleScene(scene).render()
    write_ppm_p3(tmp_path / "o.ppm", img)
    assert (tmp_path / "o.ppm").read_text() == c["ppm"]
    assert np.array_equal(read_ppm_p3(tmp_path / "o.ppm"), img)


def test_empty_scene_renders_background(workdir):
    """The reference segfaults on an object-free scene (null obj in the empty BVH leaf);
    this build defines it: every ray misses."""
    c = CLI["empty_scene_crashes_reference"]
    assert c["returncode"] == -11
    scene = Scene(text=c["config"], asset_dir=workdir)
    assert scene.n_prims == 0 and scene.desc.n_nodes == 0
    img, st = ob.OracleScene(scene).render()
    assert np.all(img == np.array([127, 63, 255], np.uint8))     # int(255*0.5), int(255*0.25), 255
    assert st.closest_rays == 16 and st.shadow_rays == 0


def test_keywords_and_state_machine(workdir):
    txt = HEADER + """
projection parallel
shadow soft
depthcueing 0.1 0.2 0.3 0.9 0.2 50 5
attlight 1 2 3 1 0.5 0.6 0.7 1 0.1 0.01
light 0 -1 0 0 1 1 1
v 0 0 -5
v 1 0 -5
v 0 1 -5
vn 0 0 2
vt 0 0
vt 1 0
vt 0 1
mtlcolor 1 1 1 1 1 1 1 1 1 0 1 1
f 1 2 3
mtlcolor 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 10 0.5 1.5
texture textures/harbor.ppm
bump textures/bumps.ppm
f 1/1 2/2 3/3
f 1/1/1 2/2/1 3/3/1
texture textures/harbor.ppm
sphere 0 0 -9 1
mtlcolor 0.1 0.2 0.3 0.4 0.5 0.6 0.7 0.8 0.9 10 0.5 1.5
f 1//1 2//1 3//1
"""
    s = Scene(text=txt, asset_dir=workdir)
    d = s.desc
    assert s.camera.parallel == 1 and s.camera.d == 4.0
    assert d.shadow_type == 1 and d.depth_cueing == 1
    assert (d.amax, d.amin, d.distmax, d.distmin) == (np.float32(0.9), np.float32(0.2), 50.0, 5.0)
    assert d.n_lights == 2 and d.lights[0].c1 == 1.0 and d.lights[1].c1 == -1.0 and d.lights[1].pos[3] == 0.0
    assert d.n_prims == 5 and d.n_textures == 1 and d.n_normalmaps == 1
    po = s.prim_object()
    by_obj = {int(po[p]): p for p in range(d.n_prims)}
    fl = lambda o: d.prim_flags[by_obj[o]]
    assert fl(0) & 2 and not (fl(0) & 4)                  # light avatar: material 1 1 1 1 1 1 1 1 1 0
    assert fl(1) & 4 and d.prim_texture[by_obj[1]] == 0 and d.prim_normalmap[by_obj[1]] == 0
    assert fl(2) & 4 and d.prim_normalmap[by_obj[2]] == -1   # `bump` is one-shot
    assert fl(3) & 1 and fl(3) & 4                        # textured sphere (texture re-used by name)
    assert not (fl(4) & 4)                                # mtlcolor switches texturing off
    n = np.ctypeslib.as_array(d.prim_normals, shape=(d.n_prims, 9))[by_obj[4]]
    assert np.allclose(n, [0, 0, 1] * 3)                  # vn is stored normalised
    # normal-map texels are remapped 2c-1 on load, colour texels are c/255
    t, m = d.textures[0], d.normalmaps[0]
    tex = np.ctypeslib.as_array(d.texels, shape=(d.n_texels, 3))
    assert tex[t.offset:t.offset + t.count].min() >= 0 and tex[m.offset:m.offset + m.count].min() < 0
    assert t.width == 512 and t.height == 256 and t.count == 512 * 256


def test_textrue_alias_and_output_name(workdir, tmp_path):
    s = Scene(text=HEADER + "textrue textures/harbor.ppm\n", asset_dir=workdir)
    assert s.desc.n_textures == 1
    for name, out in (("a.txt", "a.ppm"), ("a.b.txt.txt", "a.b.ppm"), ("noext", "noext.ppm")):
        p = tmp_path / name
        p.write_text(HEADER)
        assert Scene(p, asset_dir=workdir).output_name == str(tmp_path / out)


def test_camera_matches_reference_pixels(workdir):
    """Camera block of Renderer::render, :65-100: primary rays of the oracle camera hit exactly
    what the golden image shows (pinned by the image tests); here the closed-form pieces."""
    s = Scene(text=fixtures.bunny_shadow_config(800, 600), asset_dir=workdir)
    c = s.camera
    assert (c.width, c.height, c.parallel, c.d) == (800, 600, 0, 1.0)
    assert np.allclose(list(c.n), [0, 0, -1])
    # hfov 90 -> half width tan(45 deg) with the reference's short M_PI, in float
    wh = np.float32(np.tan(np.float32(np.float64(np.float32(45.0)) * 3.1415926535897 / 180.0)))
    assert np.float32(c.ul[0]) == np.float32(0.0) - wh
    assert np.allclose(list(c.delta_h), [2 * wh / 799, 0, 0], rtol=1e-6)
    assert np.allclose(list(c.c_off_h), [2 * wh / 1600, 0, 0], rtol=1e-6)
    s.set_imsize(3840, 2160)
    assert (s.camera.width, s.camera.height) == (3840, 2160)


def test_bunny_loader_matches_main_cpp(workdir):
    s = Scene.from_workdir(workdir, fixtures.write_config(workdir, "bl", fixtures.water_bunny_tex_config(8, 6)).stem)
    d = s.desc
    assert d.n_prims == 4970 and d.n_materials == 2
    mats = [d.materials[i] for i in range(2)]
    bunny = [m for m in mats if abs(m.eta - 1.52) < 1e-6][0]
    assert (bunny.ka, bunny.kd, bunny.ks, bunny.n, bunny.alpha) == tuple(np.float32(x) for x in (0.05, 0.1, 0.1, 64, 0.2))
    g = Scene.from_workdir(workdir, "bl", glass=True).desc
    assert any(abs(g.materials[i].eta - 1.33) < 1e-6 and g.materials[i].ks == np.float32(0.2) for i in range(2))
    # no bunny.obj in the cwd -> silently skipped (main.cpp:47)
    assert Scene.from_workdir(workdir, "bl", bunny=False).n_prims == 2


FACE_FORMS = [r"[0-9]+", r"[0-9]+//[0-9]+", r"[0-9]+/[0-9]+", r"[0-9]+/[0-9]+/[0-9]+"]    # PPMGenerator.hpp:289-292
FACE_TOKENS = ["1", "01", "1/1", "1//1", "1/1/1", "", "/", "//", "1/", "1//", "/1", "//1", "1/1/", "1//1/1", "1/1//1",
               "1/1/1/1", "a", "1a", "1/a", "-1", "+1", "1.0", "1 ", "1///1", "1/ /1", "9999999999"]


@pytest.mark.parametrize("tok", FACE_TOKENS)
def test_face_token_forms_match_reference_regexes(tok, workdir):
    """processFace accepts a face iff all three corners fullmatch the SAME one of the reference's four regexes;
    the hand-written matcher in config_parser.cpp must agree token by token."""
    import re
    if any(c.isspace() for c in tok) or tok == "":
        pytest.skip("not a single >>-token")
    form = next((i for i, f in enumerate(FACE_FORMS) if re.fullmatch(f, tok)), None)
    body = HEADER + "mtlcolor 1 1 1 1 1 1 0.2 0.6 0.2 10 1 1\nv 0 0 -3\nv 1 0 -3\nv 0 1 -3\nvn 0 0 1\nvt 0 0\n"
    txt = body + f"f {tok} {tok} {tok}\n"
    if form is None or tok in ("9999999999",):
        with pytest.raises(SceneError):
            Scene(text=txt, asset_dir=workdir)
    elif tok == "01":
        assert Scene(text=txt, asset_dir=workdir).n_prims == 1
    else:
        assert Scene(text=txt, asset_dir=workdir).n_prims == 1
        # mixing it with a different valid form is rejected like the reference's per-form triple match
        other = "1//1" if tok != "1//1" else "1"
        with pytest.raises(SceneError) as e:
            Scene(text=body + f"f {tok} {other} {tok}\n", asset_dir=workdir)
        assert "f face information is not valid" in str(e.value)
