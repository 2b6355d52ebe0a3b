"""Sweep runtime switches with short bench runs; prints hist/s per setting.
    python tools/tune.py <histories per exposure> "<refill>[,<log2 wave>]/<palette 0|1>/<l2persist 0|1>/<pipes 1|2>;..." """
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hist = sys.argv[1] if len(sys.argv) > 1 else "277778"
settings = sys.argv[2].split(";") if len(sys.argv) > 2 else ["8/1/0/2", "8/1/0/1", "8,24/1/0/2", "8/0/0/2", "4/1/0/2"]
for b in settings:
    parts = (b.split("/") + ["1", "0", "2"])[:4]
    env = dict(os.environ, DXMCB200_BATCH=parts[0], DXMCB200_PALETTE=parts[1], DXMCB200_L2PERSIST=parts[2], DXMCB200_PIPES=parts[3])
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--histories", hist, "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        km = j['roofline']['kernel_ms_per_step']
        print(f"setting {b:10s} value {j['value']:.4e} hist/s  ms/step {j['ms_per_step']:.1f}  kernel ms/step " + " ".join(f"{k}={v:.0f}" for k, v in km.items()), flush=True)
    except Exception as e:
        print("setting", b, "failed", e, out.stderr[-500:], flush=True)
