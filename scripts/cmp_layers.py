"""Compare per-launch tables written by `bench.py --layers`: usage cmp_layers.py base.json other.json [...]"""
import json, sys
tabs = [json.load(open(p)) for p in sys.argv[1:]]
names = [r["launch"] for r in tabs[0]]
print(f"{'launch':28s}" + "".join(f"{p.split('/')[-1][:14]:>15s}" for p in sys.argv[1:]))
for i, n in enumerate(names):
    print(f"{n:28s}" + "".join(f"{t[i]['ms'] * 1000:15.1f}" for t in tabs))
print(f"{'total (us)':28s}" + "".join(f"{sum(r['ms'] for r in t) * 1000:15.1f}" for t in tabs))
