import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception:
        if l.strip(): print(l[:200].rstrip())
        continue
    print(d["config"]["physics"], d.get("run", {}).get("line_search", "")[:12], "%.3fM" % (d["value"]/1e6), "%.2f ms" % d["ms_per_step"], {k: round(v,2) for k,v in d["roofline"]["kernel_ms_all"].items()}, "e2e %.3fM" % (d["e2e"]["value"]/1e6))
