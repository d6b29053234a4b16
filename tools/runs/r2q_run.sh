ncu --set full --clock-control none --import-source on -k regex:slos_thin6 -c 1 -f -o gpurun_out/r2q_thin6 python tools/profile_slos.py 12 24 1 > gpurun_out/r2q_prof.log 2>&1
ncu -i gpurun_out/r2q_thin6.ncu-rep --page raw --csv > gpurun_out/r2q_thin6_raw.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --spot 0 > gpurun_out/r2q_launch_bench.log 2>&1
tail -2 gpurun_out/r2q_prof.log; ls -la gpurun_out | grep r2q
