set -x
python tools/newton_time.py > gpurun_out/r02_p15_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_newton_inner -s 1 -c 1 -o gpurun_out/r02_inner -f python tools/newton_time.py > gpurun_out/r02_p15_ncu.log 2>&1
tail -3 gpurun_out/r02_p15_ncu.log
