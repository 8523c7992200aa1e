"""Runs tools/int8_peak (sustained full-chip tcgen05 kind::i8 rate) while sampling SM clocks and throttle reasons with
nvidia-smi, and writes gpurun_out/int8_peak.json -- the measured denominator of bench.py's roofline (copied to
profiles/int8_peak_r2.json).  gpurun only."""
import json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler  # noqa: E402

secs = sys.argv[1] if len(sys.argv) > 1 else "4"
samp = ClockSampler(0)
time.sleep(0.3)
out = subprocess.run([os.path.join(ROOT, "tools", "int8_peak"), secs], capture_output=True, text=True, timeout=120)
clocks = samp.stop()
res = json.loads(out.stdout.strip().splitlines()[-1])
res["clocks"] = clocks
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "int8_peak.json"), "w"), indent=1)
print(json.dumps(res))
