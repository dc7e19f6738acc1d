"""bench.py contract on CPU: the reference arm (the oracle port on the host cores) prints exactly one JSON line
with the keys the driver reads, and the algorithmic-byte table covers every launcher name the library profiles."""
import json
import os
import subprocess
import sys

from kcvae_testlib import ROOT


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--config", "cfg1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["global_batch"] == 16 and d["config"]["name"] == "cfg1"


def test_both_arms_describe_the_same_config():
    """`config` is built by one function for --impl ours and --impl reference: the driver's same_config check compares them."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(batch_per_gpu=0, scaling="strong")
    for name, spec in bench.CONFIGS.items():
        for world in (1, 2, 8):
            c = bench.config_dict(name, world, a)
            assert c["global_batch"] == spec["global_batch"] and c["batch_per_gpu"] * world == spec["global_batch"]
            assert c["parallelism"] == f"dp{world}" and c["name"] == name
    assert bench.local_batch("cfg2", 8, a) == 32 and bench.local_batch("cfg5", 8, a) == 64 and bench.local_batch("cfg4", 8, a) == 128
    import oracle.kcvae_oracle as O
    assert bench.score_step_bytes(O.readme_config(), 1) == 12_785_124          # 12.79 MB / frame


def test_algorithmic_byte_table_matches_the_survey_figures():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import kcvae_oracle as O
    tab, step_bytes = bench.algorithmic_bytes(O.readme_config(), 32, 1)
    # train = 31.29 MB / image + 191.1 MB / step fixed
    assert abs(step_bytes - (32 * 31_289_800 + 191_137_160)) < 32 * 2000
    for k in ("dec.tail/tc_tail_fused", "dec.out.bwd/tc_out_dgrad", "dec.out.bwd/tc_out_wgrad", "dec.convT_last.bwd/tc_convT_wgrad",
              "dec.convT_last.bwd/tc_convT_dgrad", "dec.convT.fwd/tc_convT_few_fwd", "enc.conv0.fwd/conv3x3", "enc.conv0.bwd/wgrad",
              "enc.conv1.bwd/conv3x3", "loss/image_stats", "optimizer/adam", "dec.dense.fwd/dense_wide_fwd"):
        assert tab.get(k, 0) > 0, k
    assert tab["optimizer/adam"] == 7 * 4 * 4_778_429
    # the general engine's launchers have entries too (per_kernel roofline of the default path)
    for k in ("enc.conv0.fwd/gen_conv", "enc.conv1.bwd/gen_dgrad", "enc.conv0.bwd/gen_wgrad", "dec.convT.bwd/gen_wgrad",
              "dec.dense.fwd/gen_dense", "dec.dense.bwd/gen_dense_wgrad", "dec.dense.bwd/gen_dense_dgrad"):
        assert tab.get(k, 0) > 0, k
    # 227.0 / 652.0 MFLOP per frame (README config), 15.41 / 45.98 GFLOP (scaled instance)
    assert bench.flops_per_frame(O.readme_config()) == (227_003_648, 651_980_544)
    f5 = bench.flops_per_frame(O.scaled_config())
    assert abs(f5[0] - 15.41e9) < 0.01e9 and abs(f5[1] - 45.98e9) < 0.02e9
