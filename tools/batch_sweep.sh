for b in 50 60 75 100 150; do
  python bench.py --no-cpu --batch $b --steps 8 > gpurun_out/r2ab_b$b.json 2>gpurun_out/r2ab_b$b.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2ab_b$b.json").read().strip().splitlines()[-1])
print("batch=$b value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"raw",round(d["e2e_raw"]["value"]),"ms/step",round(d["ms_per_step"],3))
PY
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
