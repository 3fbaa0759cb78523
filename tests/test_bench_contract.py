"""bench.py's driver contract, checked without a GPU through the reference arm: stdout carries exactly ONE line, it is
JSON, and it has the keys the driver reads (library banners - NCCL prints one to stdout - must not leak onto it)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, WF_CPU_GRID="1x10x13", WF_CPU_REPS_REF="1")   # a tiny sample: the contract, not the number
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["metric"] == "denoising_steps_per_sec_wan2.1_i2v_14b_480p_81f_irr_flf_dsg"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reference arm runs on OUR arm's configuration object: both arms build it with wan_config()
    assert d["config"] == {"workload": "Wan2.1-I2V-14B 480x832 81f guided sampling (IRR+FLF+DSG), 0 guided + 1 plain timed steps "
                                       "(the 15:35 mix of the 50-step run)",
                           "tokens": 32760, "dit_layers": 40, "dit_forwards_timed": 2, "vae": "fp32 storage, tf32 tensor-core convs",
                           "parallelism": "single GPU",
                           "l2_policy": "inputs larger than L2 (33 GB of weights, 0.67 GB activations streamed per GEMM)"}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"config": dict(wan_config(args, world), dit_forwards_timed=fwd - 4 * W)' in src       # run_ours
    assert '"config": wan_config(args, int(os.environ.get("WORLD_SIZE", "1")))' in src            # run_reference


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, WF_CPU_GRID="1x10x13", WF_CPU_REPS_REF="1", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_longcat_cpu_baselines_on_a_tiny_sample():
    """The LongCat lines' cpu_baseline (oracle LongCat block + VAE, extrapolated) - run in a child process on a tiny sample,
    because importing bench.py re-points file descriptor 1."""
    code = (
        "import importlib.util, json, sys, types\n"
        f"spec = importlib.util.spec_from_file_location('bench_mod', {os.path.join(ROOT, 'bench.py')!r})\n"
        "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
        "out = {}\n"
        "for refine, (h, w) in ((False, (480, 832)), (True, (704, 1280))):\n"
        "    a = types.SimpleNamespace(frames=93, height=h, width=w, steps=8, no_cpu_baseline=False)\n"
        "    out['refine' if refine else 'i2v'] = b.guarded_baseline(b.cpu_baseline_longcat, a, refine)\n"
        "a.no_cpu_baseline = True\n"
        "out['off'] = b.guarded_baseline(b.cpu_baseline_longcat, a, True)\n"
        "out['err'] = b.guarded_baseline(lambda args: 1 / 0, a.__class__(no_cpu_baseline=False))\n"
        "sys.stderr.write('RESULT ' + json.dumps(out) + '\\n')\n")
    env = dict(os.environ, WF_CPU_GRID="2x10x13", WF_CPU_REPS="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(next(l for l in r.stderr.splitlines() if l.startswith("RESULT "))[7:])
    for k, tokens in (("i2v", 37440), ("refine", 56320)):
        d = out[k]
        assert d["kind"] == "port" and d["unit"] == "steps/s" and d["value"] > 0 and d["cores"] >= 1, d
        assert f"L={tokens} tokens x 48 blocks" in d["sample"], d["sample"]
    assert "5:3 guided:plain" in out["i2v"]["sample"] and out["i2v"]["vae_round_trip_s"] > 0 and out["refine"]["vae_round_trip_s"] is None
    assert out["off"] is None and out["err"]["error"].startswith("ZeroDivisionError")
