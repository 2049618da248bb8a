import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "200", "--warmup", "20", "--no-cpu-baseline"] + sys.argv[1:], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    st = d["kernels"]["stage_ms"] if "kernels" in d else None
    print("FPS %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v, 4) for k, v in (st or {}).items()})
    for k in ("serial_ms_per_step", "serial_ms_p50", "serial_ms_p95"):
        if k in d: print(k, d[k])
except Exception as e:
    print("ERR", e, out.stdout[-2000:], out.stderr[-3000:])
