"""include/rast_draw_frame.hpp against the reference's REAL headers (headers/drawing.h:16-18, material.h:11-25 + INTEGRATION.md
section 3's one-line friend patch, light.h, face.h, arguments.h, vendored CImg): tests/shim_real_headers_main.cpp is compiled in
the dev container from where the reference lies (tests/orc.py::build_shim_real_headers), loads Suzanne with the reference's own
unmodified loader, and calls rast::draw_frame with the reference's own vectors and CImg buffers.  The binary travels to the
GPU box inside oracle/_ref/ (no reference sources do)."""
import os
import shutil
import subprocess

import pytest

import orc
import scenes as S


def _scene_dir(tmp_path):
    """Suzanne.obj + an .mtl whose map_Kd is a binary PPM: CImg reads PNG only through ImageMagick / libpng, which this image lacks."""
    from PIL import Image
    d = str(tmp_path) + "/"
    shutil.copy(os.path.join(S.DATA, "Suzanne.obj"), d + "Suzanne.obj")
    shutil.copy(os.path.join(S.DATA, "threepoint.csv"), d + "threepoint.csv")
    Image.open(os.path.join(S.DATA, "SuzanneTex.png")).convert("RGB").save(d + "SuzanneTex.ppm")
    open(d + "Suzanne.mtl", "w").write(open(os.path.join(S.DATA, "Suzanne.mtl")).read().replace("SuzanneTex.png", "SuzanneTex.ppm"))
    return d


def test_shim_compiles_against_the_reference_headers_and_fails_loudly_without_a_gpu(tmp_path):
    if not os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")):
        pytest.skip("dev-container test: /root/reference is absent")
    exe = orc.build_shim_real_headers()
    assert exe and os.path.exists(exe)
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("a GPU is present; the run is checked by the gpu-marked test")
    d = _scene_dir(tmp_path)
    p = subprocess.run([exe, d + "Suzanne.obj", d, d + "threepoint.csv", "160", "120"], capture_output=True, text=True, timeout=120)
    assert "Loading 968 triangles" in p.stdout and "Loaded texture" in p.stdout  # the reference's own loader and Material constructor ran
    assert p.returncode == 1 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_reference_types_through_the_shim_reproduce_the_reference_hashes(tmp_path):
    exe = orc.build_shim_real_headers()
    if not exe:
        pytest.skip("oracle/_ref/shim_real_headers was not built (needs /root/reference at build time)")
    case = [c for c in S.golden_cases() if c["name"] == "suzanne_160x120"][0]
    d = _scene_dir(tmp_path)
    p = subprocess.run([exe, d + "Suzanne.obj", d, d + "threepoint.csv", "160", "120"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    res = [l.split() for l in p.stdout.splitlines() if l.startswith(("RESULT", "EDITED"))]
    assert res[0][0] == "RESULT" and res[0][1] == case["frame_fnv"] and res[0][2] == case["depth_fnv"], (res, case["frame_fnv"], case["depth_fnv"])
    assert res[0][3] == "968" and res[0][4] == "3"
    assert abs(float(res[0][5])) > 0.1  # lights[0].trans_dir was written back (geometry.cpp:126)
    assert res[1][0] == "EDITED" and res[1][1] != res[0][1] and res[1][2] != res[0][2]  # vertices edited in place were re-read


@pytest.mark.gpu
def test_shim_opt_ins_page_locked_and_retained_cimg_buffers(tmp_path):
    """The two optional lines of INTEGRATION.md section 3 on the reference's own CImg buffers at 1920x1080 (large enough for the sparse
    copy): Session::pin_outputs page-locks them, Session::retained_outputs makes the third draw -- into buffers that still hold the
    second frame, not cleared by the caller -- rewrite only what changed.  First and third frame must carry the reference's hashes."""
    exe = orc.build_shim_real_headers()
    if not exe:
        pytest.skip("oracle/_ref/shim_real_headers was not built (needs /root/reference at build time)")
    case = [c for c in S.golden_cases() if c["name"] == "suzanne_1920x1080"][0]
    d = _scene_dir(tmp_path)
    for extra in ([], ["opt-ins"]):
        p = subprocess.run([exe, d + "Suzanne.obj", d, d + "threepoint.csv", "1920", "1080"] + extra, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        res = {l.split()[0]: l.split()[1:] for l in p.stdout.splitlines() if l.startswith(("RESULT", "EDITED", "AGAIN"))}
        assert res["RESULT"][0] == case["frame_fnv"] and res["RESULT"][1] == case["depth_fnv"], (extra, res)
        assert res["EDITED"][0] != res["RESULT"][0]
        assert res["AGAIN"][:2] == res["RESULT"][:2], (extra, res)
