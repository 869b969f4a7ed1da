# ncu evidence for profiles/ (run on the GPU box: bash tools/prof_cmd.sh <tag>), every ncu run after the same command
# exited 0 without ncu:
#   1. launch list of bench.py (one engine)                      -> gpurun_out/launches_<tag>.csv
#   2. time / DRAM bytes / pipe activity of EVERY launch of one pass of the bench workload (tools/prof_step.py)
#                                                                 -> gpurun_out/pass_<tag>.csv
#   3. --set full captures: the 4 encoder projections + keys GEMM + one decoder step's cell / query / vocabulary GEMMs,
#      and logmel<short>, feat_stats, feat_write, the 4 recurrences, keys_exp, attention, merge, finalise
#                                                                 -> gpurun_out/prof_<tag>_{gemm,rest}_raw.csv
TAG=${1:-r02}
set -x
rm -f gpurun_out/*.ncu-rep
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipeline 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_step.py 512 8 1 > gpurun_out/plain_prof.log 2>&1 || exit 1
# pass 2 of prof_step.py: skip the 10 launches of handle creation (9 weight splits, the E' GEMM) and the 216 of the warm-up pass
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -s 226 -c 216 --csv --log-file gpurun_out/pass_${TAG}.csv \
    python tools/prof_step.py 512 8 1 > gpurun_out/ncu_pass.log 2>&1
echo "pass metrics rc=$?"
ncu --set full --clock-control none -k regex:gemm_split_pair_kernel -s 1 -c 11 -f -o /tmp/prof_${TAG}_gemm \
    python tools/prof_step.py 512 8 1 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none -k regex:"attention_stream|beam_merge|lstm_rec|logmel|feat_write|feat_stats|keys_exp|beam_finalise" \
    -c 14 -f -o /tmp/prof_${TAG}_rest python tools/prof_step.py 512 8 1 > gpurun_out/ncu_rest.log 2>&1
echo "rest full rc=$?"
ncu -i /tmp/prof_${TAG}_gemm.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_gemm_raw.csv 2>/dev/null
ncu -i /tmp/prof_${TAG}_rest.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_rest_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out/ | tail -8
