"""Prints one summary line per bench JSON: python tools/show_bench.py gpurun_out/bench_TAG_*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    st = {k: round(v["ms_per_clip"], 1) for k, v in d["roofline"].get("stages", {}).items()}
    print(f.split("bench_")[-1][:-5], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parse/frame", round(d.get("host_parse_ms_per_frame", 0), 2),
          "cpu", round(d.get("cpu_baseline", {}).get("value", 0)), st)
