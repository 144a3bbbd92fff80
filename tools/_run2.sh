set -x
python tools/sweep_times.py --shape 500 7 300 --lanes 1 --reps 1 > gpurun_out/r02_p2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tangent_rs -s 2 -c 2 -o gpurun_out/r02_rs_k1b -f python tools/sweep_times.py --shape 500 7 300 --lanes 1 --reps 1 > gpurun_out/r02_p2_ncu.log 2>&1
tail -3 gpurun_out/r02_p2_ncu.log
