import json
import sys
for l in sys.stdin:
    if not l.startswith('{'):
        continue
    r = json.loads(l); p = r["per_rank"]
    print("seg", r.get("seg_len"), "mc", r.get("multicast"), "gpus", r["n_gpus"],
          [(x["k1_local_us"], x["k1_multicast_us"], x["barrier_us"], x["layer_us"]) for x in p])
