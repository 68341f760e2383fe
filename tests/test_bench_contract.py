"""CPU: the parts of the bench.py contract that can be checked without a GPU -- the reference arm's JSON line
(`bench.py --impl reference`: the CPU restatement of the reference graph on the host cores) and the bookkeeping
files the GPU arm reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--workload", "d0_infer_b1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["config"]["workload"] == "d0_infer_b1" and d["config"]["image_size"] == 512
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_traffic_table_points_at_committed_profiles():
    sys.path.insert(0, ROOT)
    import bench
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for wl, e in t.items():
        if wl.startswith("_"):
            continue
        assert wl in bench.WORKLOADS
        kinds = {e["kernel"]: e} if "kernel" in e else e      # round-1 layout / one entry per kernel kind
        for kind, k in kinds.items():
            assert os.path.exists(os.path.join(ROOT, k["source"])), k["source"]
            assert bench.measured_traffic(wl, kind) == k["bytes_per_launch"] > 0
        assert bench.measured_traffic(wl, "no-such-kernel") is None
    assert bench.DEFAULT_WORKLOAD == "d0_train_b32"          # BASELINE.json configs[1]
