mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in ""; do
  KWAVE_B200_LIB=$PWD/k-wave-fluid-cuda_b200/libkwave_b200$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b$v.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/b$v.log") if x.startswith("{")]
if not l: print(open("gpurun_out/b$v.log").read()[-2000:])
else:
    d=json.loads(l[-1]); print("variant=$v", "ms/step", round(d["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
