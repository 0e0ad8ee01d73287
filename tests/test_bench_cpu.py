"""bench.py contract checks that need no GPU: the reference arm prints the agreed JSON line, the CUDA arm refuses to run without a
device (there is no CPU fallback to fall into)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1  # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "4k_hdr_frames_per_sec" and d["unit"] == "frames/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "configs[1]" in d["config"]["workload"] and d["config"]["resolution"] == [3840, 2160]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "frames" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cuda_arm_refuses_to_run_without_a_device():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and r.stdout.strip() == ""


def test_headline_workload_selection():
    """configs[1] is quoted on demo map 2: bench.make_workload picks the compiled demo map where build() produced it and the seeded
    stand-in otherwise (or on request); the two arms describe whichever they use in config.workload / config.map."""
    sys.path.insert(0, ROOT)
    import bench
    from square_wheel_b200 import scene as S

    sc, is_demo = bench.make_workload("standin")
    try:
        assert not is_demo and sc.info.num_polygons > 10000 and "stand-in" in bench.workload_config(False)["workload"]
    finally:
        sc.close()
    sc, is_demo = bench.make_workload(None)
    try:
        assert is_demo == os.path.exists(S.DEMO2_MAP_PATH)
        cfg = bench.workload_config(is_demo)
        assert ("demo map 2" in cfg["workload"]) == is_demo and ("demo2.sqwm" in cfg["map"])
        poses = sc.camera_path(bench.NUM_POSES)
        assert len(poses) == bench.NUM_POSES
        if is_demo:
            # the path starts at the demo's info_player_start (2560 -648 -352, eye 40 above) and stays inside the map's bounds
            assert abs(poses[0][0] - 2560) < 600 and abs(poses[0][1] + 648) < 600  # (the first pose is half a step along the path)
    finally:
        sc.close()
