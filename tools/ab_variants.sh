for v in "$@"; do
  cp variants/lib_$v.so chinese_asr_b200/libasr_b200.so
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/ab_$v.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/ab_$v.log") if x.startswith("{")][-1]
d=json.loads(l)
print("$v", "ms/step", round(d["ms_per_step"],3), "stages", {k:round(x,3) for k,x in d["stage_ms"].items() if k in ("enc_recurrence","enc_input_gemm","features")}, d["clocks"]["sm_mhz"])
PY
done
