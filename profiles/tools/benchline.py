import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["ms_per_launch"].items()}, "e2e", round(d["e2e"]["value"]))
