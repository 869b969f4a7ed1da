# ncu --set full captures of the decoder-step kernels at the bench workload (512 x 10 s, bw=8), one launch each,
# after the same command exited 0 without ncu.  Usage (on the GPU box): bash tools/prof_cmd.sh <tag>
TAG=${1:-r02}
set -x
python tools/prof_step.py 512 8 1 > gpurun_out/plain_prof.log 2>&1 || exit 1
for spec in "vocab gemm_split_pair_kernel 11" "cell gemm_split_pair_kernel 9" "merge beam_merge_kernel 1" "attn attention_stream_kernel 1"; do
    set -- $spec
    ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 \
        python tools/prof_step.py 512 8 1 > gpurun_out/ncu_$1.log 2>&1
    echo "$1 rc=$?"
done
ls -la gpurun_out/*.ncu-rep
