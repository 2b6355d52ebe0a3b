"""Sweep the warp regrouping thresholds (DXMCB200_BATCH=<refill>,<interact>[,<log2 chunk>]) and the palette switch
(DXMCB200_PALETTE) with short bench runs; prints hist/s per setting.
    python tools/tune.py <histories per exposure> "<batch>[/<palette>];..." """
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hist = sys.argv[1] if len(sys.argv) > 1 else "277778"
settings = sys.argv[2].split(";") if len(sys.argv) > 2 else ["4,8/1", "4,8/0", "2,8/1", "8,8/1", "4,12/1", "4,16/1", "8,16/1", "8,16/0"]
for b in settings:
    batch, _, pal = b.partition("/")
    env = dict(os.environ, DXMCB200_BATCH=batch, DXMCB200_PALETTE=pal or "1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--histories", hist, "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"],
                         env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"batch {b:10s} value {j['value']:.4e} hist/s  kernel_ms/step {j['roofline']['kernel_ms_per_step']:.1f}", flush=True)
    except Exception as e:
        print("batch", b, "failed", e, out.stderr[-500:], flush=True)
