# refresh of the launch list after the feature / split fusion + set-full capture of the changed feat_write_kernel
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_step.py 512 8 1 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none -k regex:feat_write_kernel -s 1 -c 1 -f -o /tmp/prof_fw python tools/prof_step.py 512 8 1 > gpurun_out/ncu_prof_fw.log 2>&1
echo "set full rc=$?"
ncu -i /tmp/prof_fw.ncu-rep --page raw --csv > gpurun_out/prof_r01_fw_raw.csv 2>/dev/null
du -sh gpurun_out
