"""Top stall-sample instructions per kernel from `ncu -i <rep> --page source --csv` (harness, not product code).
Usage: ncu -i gpurun_out/prof.ncu-rep --page source --csv | python tools/ncu_source_stalls.py [top_n] > profiles/ncu_source_rNN_top_stalls.txt"""
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows = list(csv.reader(sys.stdin))
kernels, cur = [], None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        cur = {"name": r[1], "ins": []}
        kernels.append(cur)
    elif r[0] == "Address":
        cur["col"] = {n: i for i, n in enumerate(r)}
    elif cur is not None and "col" in cur and r[0].startswith("0x"):
        c = cur["col"]
        try:
            cur["ins"].append((int(r[c["Warp Stall Sampling (All Samples)"]]), r[c["Source"]].strip()))
        except (ValueError, KeyError, IndexError):
            pass
print("# ncu --page source (SASS, warp-stall sampling): the instructions with the most stall samples per kernel")
seen = set()
for k in kernels:
    key = (k["name"], tuple(k["ins"][:50]))
    if key in seen:                       # the report lists every launch once per source view
        continue
    seen.add(key)
    total = sum(n for n, _ in k["ins"]) or 1
    short = k["name"].split("(const")[0].split("(unsigned")[0]
    print(f"\n== {short}  (total samples {total}, {len(k['ins'])} SASS instructions)")
    for n, src in sorted(k["ins"], key=lambda t: -t[0])[:top_n]:
        print(f"{100.0 * n / total:5.1f}%  {src}")
