#!/bin/bash
# Build libasr_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [--clean] [extra nvcc flags]
# One nvcc -c per source, in parallel; an object is rebuilt when its source, a shared header, the public
# header or this script is newer.  --clean (what __graft_entry__.build() passes) or extra flags force a full
# rebuild.  build.log records, per object, whether it was rebuilt and the ptxas resource usage (-Xptxas -v).
set -e
cd "$(dirname "$0")"
SRC=chinese_asr_b200/csrc
OUT=chinese_asr_b200/libasr_b200.so
OBJ=build/obj
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
SOURCES="api features gemm_tc encoder encoder_tc3 decoder wer resample"
if [ "$1" = "--clean" ]; then shift; rm -rf $OBJ; fi
mkdir -p $OBJ
echo "# build.sh $(date -u +%Y-%m-%dT%H:%M:%SZ)  $($NVCC --version | tail -1)" > build.log
echo "# flags: $FLAGS $*" >> build.log
pids=()
names=()
objs=""
for f in $SOURCES; do
    o=$OBJ/$f.o
    objs="$objs $o"
    stale=0
    if [ ! -f $o ] || [ $# -gt 0 ]; then stale=1; else
        for d in $SRC/$f.cu $SRC/*.cuh include/asr_b200.h build.sh; do
            [ $d -nt $o ] && stale=1
        done
    fi
    if [ $stale = 1 ]; then
        ( $NVCC $FLAGS "$@" -c $SRC/$f.cu -o $o > $OBJ/$f.log 2>&1 ) &
        pids+=($!)
        names+=($f)
    else
        echo "== $f.cu: up to date ($o)" >> build.log
    fi
done
fail=0
for i in "${!pids[@]}"; do
    if ! wait ${pids[$i]}; then fail=1; echo "nvcc failed on ${names[$i]}.cu"; fi
    echo "== ${names[$i]}.cu: rebuilt" >> build.log
    cat $OBJ/${names[$i]}.log >> build.log
done
if [ $fail = 1 ]; then rm -f $(for n in "${names[@]}"; do echo $OBJ/$n.o; done); grep -E "error" build.log | head -40; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $objs -o $OUT >> build.log 2>&1 || { tail -20 build.log; exit 1; }
echo "== linked $OUT ($(stat -c %s $OUT) bytes) from:$objs" >> build.log
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
grep -E "bytes spill stores" build.log | grep -v " 0 bytes spill stores" | sort | uniq -c | head -10 || true
echo "built $OUT (${#names[@]} objects rebuilt)"
