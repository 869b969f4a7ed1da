# ncu evidence for profiles/: launch list of bench.py + one --set full capture of a pass (encoder + 2 decoder steps).
# The .ncu-rep stays on the box (/tmp: gpurun_out/ is capped at 64 MiB); its raw page comes back as CSV.
set -x
rm -f gpurun_out/*.ncu-rep
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_step.py 512 8 1 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --launch-skip 258 --launch-count 28 -f -o /tmp/prof_r01_all python tools/prof_step.py 512 8 1 > gpurun_out/ncu_prof.log 2>&1
echo "set full rc=$?"
ncu -i /tmp/prof_r01_all.ncu-rep --page raw --csv > gpurun_out/prof_r01_all_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out/ | tail -12
