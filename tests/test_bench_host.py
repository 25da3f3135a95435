"""bench.py host logic that needs no GPU: workload models, config strings, the reference arm."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_work_models_match_survey_table():
    # flops / bytes per element of every BASELINE config
    want = {"div_p4": (7980, 1192), "grad_p4": (7980, 1192), "lift_p4": (17040, 3072),
            "tp_p7": (8192, 8192), "wave_p4": (33000, 5384), "wave_p4_f32": (33000, 2692),
            "div_p4_f32": (7980, 596), "lift_p4_f32": (17040, 1536)}
    for name, (fl, by) in want.items():
        w = bench.Work(name)
        assert (w.flops, w.bytes) == (fl, by), name


def test_every_baseline_config_is_in_the_default_suite():
    names = {n for n, _, _ in bench.SUITE} | {"div_p4"}
    assert {"grad_p4", "div_p4", "lift_p4", "wave_p4", "wave_p4_f32", "tp_p7",
            "grad_p4_f32", "div_p4_f32", "lift_p4_f32"} <= names
    assert ("grad_p4", 100_000, False) in bench.SUITE          # BASELINE configs[0] size
    assert set(bench.STRONG_SUITE) == {"div_p4", "wave_p4", "tp_p7"}


def test_config_strings_are_per_workload():
    w = bench.Work("lift_p4")
    cfg = bench.make_config(w, 4_000_000, 8, "weak", False)
    assert "12.29 GB" in cfg["l2_policy"] and "configs[2]" in cfg["workload"]
    assert cfg["elements_total"] == 32_000_000
    assert "L2-resident" in bench.Work("grad_p4").l2_policy(100_000, False)
    assert "flushed" in bench.Work("grad_p4").l2_policy(100_000, True)


def test_shapes_of_multi_einsum_workloads():
    ins, outs = bench.Work("wave_p4").shapes(10)
    assert ins["J"] == (3, 3, 10) and outs["grad_out"] == (3, 10, 35) and len(outs) == 6


@pytest.mark.parametrize("workload", ["div_p4", "wave_p4_f32"])
def test_reference_arm_prints_one_json_line(workload):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          workload, "--elements", "2000", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, check=True, cwd=ROOT).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "GFLOP/s"
    assert line["config"]["reference_sample_elements_per_step"] == 2000 == line["config"]["elements_per_gpu"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--elements", "1000", "--steps", "1"], capture_output=True, text=True, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_sample_size_is_bounded_and_stated():
    w = bench.Work("tp_p7")
    n = bench.reference_sample_elements(w, 4_000_000)
    assert 100_000 <= n <= 4_000_000 and w.flops * n <= 4e10 * 2
