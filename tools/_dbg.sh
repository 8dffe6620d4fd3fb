mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r1l_launches.csv python tools/prof_driver.py --edge 150 --iters 64 > gpurun_out/r1l_prof.log 2>&1
tail -4 gpurun_out/r1l_prof.log
