"""CPU checks of bench.py's reference arm (the one leg that runs without a GPU) and of the tuning-switch table:
the JSON line carries every key the measurement contract names, and every dg_set_tuning key is documented."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "the reference arm prints exactly one line"
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_every_tuning_key_is_documented():
    hdr = open(os.path.join(ROOT, "include", "downgan_b200.h")).read()
    n = int(re.search(r"#define DG_TUNE_KEYS (\d+)", hdr).group(1))
    doc = hdr[hdr.index("key 0:"):hdr.index("#define DG_TUNE_KEYS")]
    for k in range(n):
        assert re.search(rf"key {k}\b", doc), f"dg_set_tuning key {k} is not described in the header"
    src = open(os.path.join(ROOT, "downgan_b200", "csrc", "dg_kernels.cu")).read()
    init = re.search(r"g_tune\[DG_TUNE_KEYS\] = \{([^}]*)\}", src).group(1)
    assert len([x for x in init.split(",") if x.strip()]) == n, "default table and DG_TUNE_KEYS disagree"
