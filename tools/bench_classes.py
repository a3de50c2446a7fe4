import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), {k:(round(v["ms"],1), v["launches"]) for k,v in d["roofline"]["classes"].items() if v["ms"]>1})
