import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
import sys, json
src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin   # file argument or stdin
for l in src:
    if l.startswith('{'):
        d = json.loads(l); r = d.get("roofline") or {}
        print("value %.0f links/s  %.2f ms/step" % (d["value"], d["ms_per_step"]))
        if d.get("e2e"): print("e2e %.0f links/s %.1f ms/step" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
        if r: print("gather %.0f GB/s frac %.3f avg %.3f ms; shares %s; path frac %.3f" % (r["achieved"], r["frac"], r["avg_launch_ms"], {k: round(v, 3) for k, v in r["share_of_step"].items()}, r["path"]["frac"]))
        if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
        print("clocks", d.get("clocks"), "host_enqueue_ms", d.get("host_enqueue_ms_per_step"), "step_ms", d.get("step_ms"))
