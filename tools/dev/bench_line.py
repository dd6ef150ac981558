import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d["ms_per_step"], round(d["value"]), [ (s["kernel"].split()[0], s["ms"]) for s in d["stages"]])
