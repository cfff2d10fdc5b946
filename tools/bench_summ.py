import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"],4), d["roofline"]["kernel_ms_per_step"], (d.get("parity") or {}).get("bitwise_equal"), 'e2e', round(d['e2e']['ms_per_step'],3))
    except Exception as e: print(f, "ERR", e)
