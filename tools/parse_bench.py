import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), {k:(round(v["ms"],3), round(v["gbs"])) for k,v in d["roofline"]["kernels"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "ERR", e)
