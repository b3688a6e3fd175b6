"""Stages tools/r2ax_case/ (git-ignored, travels with the gpurun snapshot): three hard-shadow configs, their assets and the
PPMs the UNMODIFIED reference executable (oracle/_ref/whitted_ref, needs /root/reference at build time) writes for them —
what tools/gpu_r2ax.sh compares the drop-in executable's output with on the GPU box, byte for byte, without Python."""
import shutil
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from whittedstyle_raytracer_b200 import fixtures  # noqa: E402

wd = REPO / "tools" / "r2ax_case"
wd.mkdir(exist_ok=True)
fixtures.ensure_assets(wd)
fixtures.write_config(wd, "glass_bunny", fixtures.water_bunny_tex_config(640, 400))
fixtures.write_config(wd, "hard_bunny", fixtures.bunny_shadow_config(640, 400))
# the row of glass spheres of tests/test_gpu_parity.py::test_hard_shadow_product_has_the_reference_association
text = fixtures._CAMERA.format(w=320, h=240) + "light 8 0.2 -2 1 1 1 1\n"
for k, a in enumerate([0.1, 0.25, 0.4, 0.55, 0.7, 0.85]):
    text += f"mtlcolor 0.8 0.8 0.9 1 1 1 0.2 0.6 0.3 20 {a} 1.3\nsphere {-3 + k} 0 -2 0.42\n"
text += "mtlcolor 0.7 0.7 0.7 1 1 1 0.2 0.8 0.0 10 1 1\nv -12 -0.6 6\nv 12 -0.6 6\nv 12 -0.6 -14\nv -12 -0.6 -14\nf 1 2 3\nf 1 3 4\n"
fixtures.write_config(wd, "glass_row", text)
for n in ("glass_bunny", "hard_bunny", "glass_row"):
    p = subprocess.run([str(REPO / "oracle" / "_ref" / "whitted_ref"), n + ".txt"], cwd=wd, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-500:] + p.stderr[-500:]
    shutil.move(wd / (n + ".ppm"), wd / ("expected_" + n + ".ppm"))
    print("staged", n)
