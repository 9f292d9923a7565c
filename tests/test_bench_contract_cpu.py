"""bench.py contract, CPU side: the reference arm (the CPU port, rank 0 only) prints ONE JSON line whose `config` equals what
the GPU arm prints for the same workload, whose `steps` are the whole frames it really timed, and which carries the keys the
driver reads; `--workload train --impl reference` reports that there is nothing to run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return [json.loads(l) for l in lines]


def test_reference_arm_line():
    lines = _run("--impl", "reference", "--workload", "tiny", "--steps", "3", "--warmup", "1")
    assert len(lines) == 1
    l = lines[0]
    assert l["impl"] == "reference" and l["metric"] == "fusion_layer_frames_per_sec" and l["unit"] == "frames/s"
    assert l["higher_is_better"] is True and l["vs_baseline"] is None
    assert l["steps"] == 3 and l["steps_requested"] == 3 and l["warmup"] == 1
    # steps x ms_per_step is CPU time that was really spent: value = frames / that time, one step = one whole frame
    assert abs(l["value"] - 1e3 / l["ms_per_step"]) <= 1e-2 * l["value"]
    assert l["cpu_baseline"]["kind"] == "port" and l["cpu_baseline"]["cores"] >= 1 and "nothing extrapolated" in l["cpu_baseline"]["sample"]
    assert l["e2e"] == {"value": l["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the identical config dict in both arms: only the workload label (arm-specific settings live under "run")
    sys.path.insert(0, ROOT)
    import bench
    import dcf_b200 as dcf
    wl = dcf.synthetic.make_workload("tiny", seed=100)
    assert l["config"] == {"workload": bench.workload_label("tiny", wl)}


def test_reference_arm_is_silent_on_other_ranks():
    assert _run("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_train_workload_has_no_reference():
    lines = _run("--workload", "train", "--impl", "reference")
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and "unavailable" in lines[0]
