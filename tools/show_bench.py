"""Print the headline numbers of a bench.py JSON line."""
import json
import sys

j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.0f  e2e %.0f  ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
print({k: v["ms_total"] for k, v in j["kernels"].items()})
