# A/B of library variants on one box: bash tools/ab_variants.sh <pipeline engines> <variant> ...   (variants/lib_<variant>.so)
P=$1; shift
for v in "$@"; do
  cp variants/lib_$v.so chinese_asr_b200/libasr_b200.so
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --pipeline $P > gpurun_out/ab_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/ab_$v.log") if x.startswith("{")][-1]
d=json.loads(l)
print("$v", "P=$P ms/step", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["stage_ms"].items()}, d["clocks"]["sm_mhz"])
PY
done
