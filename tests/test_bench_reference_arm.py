"""CPU: `bench.py --impl reference` (the arm the driver times beside ours) runs without a GPU and prints one JSON
line with the contract's keys.  It executes the compiled reference coder (oracle/_ref) or, without it, the C port."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
  d = json.loads(line)
  assert d["impl"] == "reference" and d["unit"] == "Msymbols/s" and d["higher_is_better"] is True
  assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
  assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
  assert "workload" in d["config"]
