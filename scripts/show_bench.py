"""Print the interesting parts of a bench.py JSON line: python scripts/show_bench.py FILE.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
for k in ["metric", "value", "ms_per_step", "encode_ms", "decode_ms", "n_gpus"]:
    print(k, d.get(k))
if d.get("roofline"):
    r = d["roofline"]
    print("roof", {k: r.get(k) for k in ["achieved", "peak", "frac", "executed_frac", "ms_per_launch", "traffic"]})
    print("dec", r.get("decode"))
print("e2e", d.get("e2e"))
print("cpu", d.get("cpu_baseline"))
print("eager", d.get("eager_gpu_baseline"))
print("module", d.get("module"))
print("sustained", d.get("sustained"))
print("collective", d.get("collective"))
print("clocks", d.get("clocks"))
print("config", d.get("config"))
for k, v in (d.get("secondary") or {}).items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()
              if a in ("ms_per_step", "search_ms", "decode_ms", "search_frac", "decode_frac", "search_kernel", "value", "error")})
