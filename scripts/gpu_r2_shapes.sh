#!/bin/bash
# BASELINE configs 3 and 4 (RealEstate10K 384x256 N=64 B=4, KITTI 768x256 N=32 B=4), both arms, 1 GPU
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
for shape in realestate kitti; do
  timeout 600 python bench.py --shape $shape --steps 10 --warmup 3 --no-render > gpurun_out/bench_${shape}_ours.json 2> gpurun_out/bench_${shape}_ours.err; echo "$shape ours rc=$?"
  timeout 900 python bench.py --impl reference --shape $shape --steps 10 --warmup 3 --no-render > gpurun_out/bench_${shape}_ref.json 2> gpurun_out/bench_${shape}_ref.err; echo "$shape ref rc=$?"
  python - $shape <<'PY'
import json, sys
shape = sys.argv[1]
for arm in ("ours", "ref"):
    f = "gpurun_out/bench_%s_%s" % (shape, arm)
    try:
        d = json.loads([l for l in open(f + ".json").read().strip().splitlines() if l.startswith("{")][-1])
        extra = ""
        if "fast" in d: extra = " | bf16 %.1f img/s" % d["fast"].get("value", -1)
        print("%-10s %-5s %.2f img/s  %.1f ms/step  e2e %.2f%s" % (shape, arm, d.get("value", -1), d.get("ms_per_step", -1), d.get("e2e", {}).get("value", -1), extra), d.get("unavailable", ""))
    except Exception as e:
        print(shape, arm, "no result:", e); print(open(f + ".err").read()[-1200:])
PY
done
