#!/bin/bash
# ncu --set full of the row-split cluster kernels on the large grids; only the text summaries travel back
set -x
run() {  # tag, shape..., lanes
  tag=$1; shift
  timeout 600 ncu --set full --clock-control none -k 'regex:k_(backward|forward)_tangent' --launch-skip 2 --launch-count 2 -f -o /tmp/$tag python tools/sweep_times.py "$@" --reps 1 > gpurun_out/r02c_$tag.log 2>&1
  echo "#### $tag: python tools/sweep_times.py $*" >> gpurun_out/r02c_rowsplit_ncu_summary.txt
  python tools/ncu_summary.py /tmp/$tag.ncu-rep >> gpurun_out/r02c_rowsplit_ncu_summary.txt 2>&1
}
: > gpurun_out/r02c_rowsplit_ncu_summary.txt
run c4_k64 --shape 2000 11 500 --lanes 64
run c5_k64 --shape 1000 7 300 --lanes 64
run c5_k444 --shape 1000 7 300 --lanes 444
run c2_k1 --shape 500 7 300 --lanes 1
cat gpurun_out/r02c_rowsplit_ncu_summary.txt
