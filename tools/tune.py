"""Sweep the warp regrouping thresholds (DXMCB200_BATCH) with short bench runs; prints hist/s per setting."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hist = sys.argv[1] if len(sys.argv) > 1 else "277778"
for b in ["8,8", "4,8", "8,12", "8,16", "12,16", "16,16", "16,24"]:
    env = dict(os.environ, DXMCB200_BATCH=b)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--histories", hist, "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"batch {b:6s} value {j['value']:.4e} hist/s  kernel_ms/step {j['roofline']['kernel_ms_per_step']:.1f}", flush=True)
    except Exception as e:
        print("batch", b, "failed", e, out.stderr[-500:], flush=True)
