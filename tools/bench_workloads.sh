#!/bin/bash
# Runs bench.py (our arm, no CPU baseline) for each named workload and prints the stage times.
cd "$(dirname "$0")/.."
for w in "$@"; do
  t0=$(date +%s.%N)
  out=$(timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-extra --workload $w 2>&1 | tail -1)
  t1=$(date +%s.%N)
  python - "$w" "$out" "$t0" "$t1" <<'PY'
import json, sys
w, out = sys.argv[1], sys.argv[2]
try:
    j = json.loads(out)
    st = j["roofline"]["stages"]; c = dict(j["config"]); c.update(j.get("details", {}))
    print("%-18s %8.1f us %6.2f Gtri/s | tris %d draws %d surv %d | " % (w, j["ms_per_step"] * 1e3, j["value"] / 1e9, c["triangles_per_frame"], c["draws_rank0"], c["surviving_triangles_rank0"])
          + " ".join("%s %.1f(%.0f%%)" % (k, st[k]["ms"] * 1e3, 100 * st[k].get("frac", 0)) for k in ("prep", "vertex", "cull", "sort", "emit"))
          + " | e2e %.2f ms (h2d %.1f MB d2h %.1f MB) | wall %.0f s" % (j["e2e"]["ms_per_step"], j["e2e"]["h2d_bytes_per_step"] / 1e6, j["e2e"]["d2h_bytes_per_step"] / 1e6, float(sys.argv[4]) - float(sys.argv[3])))
except Exception as e:
    print(w, "FAILED", out[-400:])
PY
done
