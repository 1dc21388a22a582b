S=$SECONDS
python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err
echo "bench wall $((SECONDS-S)) s rc $?"
S=$SECONDS
python bench.py --impl reference > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err
echo "reference wall $((SECONDS-S)) s rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_default.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"])
print([(k, v.get("value"), v.get("ms_per_step")) for k,v in d["secondary"].items()])
r=json.load(open("gpurun_out/final_reference.json")); print("ref", r["value"], r["steps"], r["ms_per_step"])
PY
