"""The closed form of `k` rounded float additions used by the shadow-map depth kernels (csrc/depth_kernels.cu: f32_add_repeated)
against the plain loop of the reference (rasterizer.rs:1463 `inv_z += d_inv_z_dx`, :1470 `line_inv_z += d_inv_z_dy`).
tools/f32_add_repeated_check.c holds the host copy of the algorithm and the adversarial generators (ties, binade edges, zero
crossings, denormals, random bit patterns); 3.2e8 cases were run once, the test keeps 3e6 of them."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_closed_form_equals_the_addition_loop(tmp_path):
    exe = tmp_path / "f32_add_repeated_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "tools", "f32_add_repeated_check.c"), "-lm"], check=True)
    for seed in ("1", "77"):
        r = subprocess.run([str(exe), "1500000", seed], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout
        assert " 0 bad" in r.stdout
