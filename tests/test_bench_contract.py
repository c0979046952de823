"""bench.py's reference arm (`--impl reference`: the staged Python reference and the C restatement on the host cores) runs
without a GPU, so its JSON contract is checked here; the GPU arm prints the same keys (profiles/r01_bench_*.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                           "--warmup", "1", "--cpu-envs", "64", *args], capture_output=True, text=True, env=env,
                          timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_steps_per_sec" and d["unit"] == "agent-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("c4:") and d["config"]["num_drones"] == 32
    cb = d["cpu_baseline"]
    # the unmodified Python reference where oracle/_ref (or /root/reference) is present, else the C port alone
    assert d["reference_kind"] in ("python_reference", "c_port")
    assert cb["kind"] == ("reference" if d["reference_kind"] == "python_reference" else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["port"]["kind"] == "port" and d["port"]["value"] > 0
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, ("--gpus", "2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_follows_the_workload_flag():
    r = _run(args=("--workload", "c2"))
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["config"]["num_drones"] == 8 and d["config"]["num_obstacles"] == 4
