set -x
python tools/sweep_times.py --shape 1000 7 300 --lanes 64 --reps 1 > gpurun_out/r02_p9_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tangent_rs -s 2 -c 2 -o gpurun_out/r02_rs_c5 -f python tools/sweep_times.py --shape 1000 7 300 --lanes 64 --reps 1 > gpurun_out/r02_p9_ncu.log 2>&1
tail -3 gpurun_out/r02_p9_ncu.log
