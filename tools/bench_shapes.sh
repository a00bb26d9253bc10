#!/bin/bash
# One bench line per named BASELINE shape (1 GPU); appended to gpurun_out/shapes.jsonl
mkdir -p gpurun_out
: > gpurun_out/shapes.jsonl
for w in cora pubmed ppi pattern; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 >> gpurun_out/shapes.jsonl 2>> gpurun_out/shapes.err
done
python - <<'PY'
import json
for line in open("gpurun_out/shapes.jsonl"):
    d = json.loads(line)
    cb = d.get("cpu_baseline") or {}
    print(d["config"]["graph"], "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"],
          "cpu %.3e (%s cores)" % (cb.get("value", 0), cb.get("cores")), "launches", d["gpu_launches"])
PY
