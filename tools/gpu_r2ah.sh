#!/bin/bash
# NOT RUN TO COMPLETION in round 2: the first attempt selected the kernels with -k regex on the template arguments, but ncu
# matches -k against the function NAME unless --kernel-name-base demangled is given (fixed below); no GPU budget was left to repeat it.
# round 2: ncu --set full captures of the group-skipping pass B (float64) and of the f32 KDE d=1 kernel with the MUFU offload
set -x
mkdir -p gpurun_out
WHICH=gskip timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"pair_kernel<double, 4, 1, 0, 0, 1>" -s 1 -c 1 -f -o gpurun_out/r2_prof_gskip python tools/profile_r2_new.py > gpurun_out/ncu_gskip.log 2>&1; tail -3 gpurun_out/ncu_gskip.log
WHICH=soft timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"pair_kernel<float, 1," -s 1 -c 1 -f -o gpurun_out/r2_prof_soft python tools/profile_r2_new.py > gpurun_out/ncu_soft.log 2>&1; tail -3 gpurun_out/ncu_soft.log
ls -la gpurun_out/*.ncu-rep
