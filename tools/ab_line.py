"""One line per bench log: python tools/ab_line.py <log> ..."""
import json
import sys

for f in sys.argv[1:]:
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "no result", open(f).read()[-800:])
        continue
    j = json.loads(l[-1])
    print("%s: %.2f Gvox/s  step %.3f ms  %s  frac %.3f" % (f, j["value"], j["ms_per_step"], j.get("kernel_ms"), j["roofline"]["frac"]))
