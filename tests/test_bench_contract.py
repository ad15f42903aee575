"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the
reference algorithm on the host cores) prints ONE JSON line with the contract's keys, ranks other than 0 stay silent,
and the algorithmic-byte accounting matches SURVEY.md section 8-d / DESIGN.md section 3."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                        "--steps", "3", "--warmup", "1", *flags], capture_output=True, text=True, cwd=ROOT, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_contract_line():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ant-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ant-steps/sec") and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0   # W >= 3
    assert d["config"]["workload"].startswith("cfg2") and d["data"] == "synthetic" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    sys.path.insert(0, ROOT)
    from oracle import ref_harness
    # the unmodified reference (checkout, or oracle/_ref built by oracle/build_ref.py) wherever it exists
    assert cb["kind"] == ("reference" if ref_harness.reference_available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "ant-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert out.strip() == ""


def test_algorithmic_bytes_follow_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WORKLOADS["cfg4"]
    E, N, P = 512, wl["n_ants"], wl["n_phero"]
    # perception: obs f32 + agent_state + reward out, 49 samples x (phero, food, walls, explored r/w), per-ant state
    per_ant = 49 * 7 * 4 + 16 + 49 * (8 * P + 8 + 3) + 82
    assert per_ant == 2793
    assert bench.algorithmic_bytes("perceive", wl, E, 7, {}) == E * N * per_ant
    assert bench.algorithmic_bytes("perceive", bench.WORKLOADS["cfg3"], 1, 6, {}) == 256 * 2597
    assert bench.algorithmic_bytes("evaporate", wl, E, 7, {"evap_mode": "lazy"}) == 0
    assert bench.algorithmic_bytes("evaporate", wl, E, 7, {"evap_mode": "dense"}) == E * 1024 * 1024 * (16 * P + 1)
    step = sum(bench.algorithmic_bytes(f, wl, E, 7, {"evap_mode": "lazy"})
               for f in ("move", "perceive", "collide", "rocks", "deposit")) / (E * N)
    assert abs(step - 3109) < 1                                              # bytes per ant-step in the bench line
    # the block-per-environment kernels carry the algorithmic bytes of the flat kernels they replace
    fused = sum(bench.algorithmic_bytes(f, wl, E, 7, {"evap_mode": "lazy"}) for f in ("perceive", "env_update_move")) / (E * N)
    assert abs(fused - step) < 1e-9
    peak, src = bench.load_peaks()
    assert 5000 < peak < 9000 and ("measured" in src or "fallback" in src)


def test_hand_typed_multi_gpu_run_becomes_the_torchrun_launch():
    sys.path.insert(0, ROOT)
    import bench
    cmd = bench.torchrun_command(4, ["--gpus", "4", "--steps", "50"], port=29777)
    assert cmd[:3] == [sys.executable, "-m", "torch.distributed.run"]
    assert cmd[3:10] == ["--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1", "--master-port", "29777"]
    assert cmd[10] == os.path.join(ROOT, "bench.py") and cmd[11:] == ["--gpus", "4", "--steps", "50"]
