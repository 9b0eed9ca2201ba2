#!/bin/bash
# usage (on the GPU box, via gpurun): scripts/gpu_check.sh <outdir> [profile]
# full -m gpu suite, solver phase times, (optional) warm launch lists of the ORORA pair, the N=1 bench line
out=$1; mkdir -p "$out"
python -m pytest tests -m gpu -q > "$out/gputest.log" 2>&1; echo rc=$? >> "$out/gputest.log"; tail -6 "$out/gputest.log"
python scripts/dbg_solver.py 2>&1 | tail -3
if [ "$2" = "profile" ]; then
  for m in cen2019 cen2018; do
    NRS_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv \
        --log-file "$out/launches_$m.csv" python scripts/prof_orora.py $m mulran > "$out/run_$m.log" 2>&1
    tail -1 "$out/run_$m.log" | cut -c1-200
    python scripts/ncu_summarize.py list "$out/launches_$m.csv" "$out/launch_list_$m.json" \
        "$m: two steady-state scan pairs through nrs_orora_pair, NRS_NO_GRAPHS=1, ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none (warm caches, serialised launches)"
    python - "$out/launch_list_$m.json" $m <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], d["total_us"] / 2, "us per pair (sum of kernel durations)")
for k in d["kernels"][:14]:
    print("%-44s x%-3d %8.2f us  share %.3f" % (k["kernel"][:44], k["launches"], k["avg_us"], k["share"]))
PY
  done
fi
python bench.py --steps 20 --warmup 5 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; tail -3 "$out/bench_n1.err"
python - "$out/bench_n1.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("SC", d["value"], "q/s", d["ms_per_step"], "ms/step, e2e", d["e2e"]["value"], "single", d["single_query"]["device_us"], "us")
o = d["orora"]
print({k: o[k] for k in ("value", "ms_per_pair", "stage_ms", "gpu_launches_per_pair")}, "separate", o["separate_calls"]["ms_per_pair"],
      "cen2018", o["cen2018"]["ms_per_pair"], "pipelined", o["pipelined"]["value"])
print("oxford", d["orora_oxford"]["value"], "stream", d["stream"]["value"], "kdtree", d["kdtree_mode"]["gpu_value"])
PY
