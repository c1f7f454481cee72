#!/bin/bash
# Runs bench.py once per libt3dgeo.so variant under build/variants/ (LD_LIBRARY_PATH override) and prints the stage times.
# usage: tools/bench_variants.sh [variant ...]   (default: all)
cd "$(dirname "$0")/.."
vs="$@"; [ -z "$vs" ] && vs=$(ls build/variants)
for v in $vs; do
  out=$(LD_LIBRARY_PATH=$PWD/build/variants/$v:$LD_LIBRARY_PATH timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra ${BENCH_ARGS} 2>&1 | tail -1)
  python - "$v" "$out" <<'PY'
import json, sys
v, out = sys.argv[1], sys.argv[2]
try:
    j = json.loads(out)
    st = j["roofline"]["stages"]
    print("%-10s %.1f us | " % (v, j["ms_per_step"] * 1e3) + " ".join("%s %.1f" % (k, st[k]["ms"] * 1e3) for k in ("prep", "vertex", "cull", "sort", "emit")) + " | e2e %.2f ms" % j["e2e"]["ms_per_step"])
except Exception as e:
    print(v, "FAILED", out[-300:])
PY
done
