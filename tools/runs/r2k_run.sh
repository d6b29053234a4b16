python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 300 gpurun_out/r2k_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --spot 0 > gpurun_out/r2k_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:slos_thin6 -c 1 -f -o gpurun_out/r2k_thin6 python tools/profile_slos.py 12 24 1 > gpurun_out/r2k_prof.log 2>&1
ncu -i gpurun_out/r2k_thin6.ncu-rep --page raw --csv > gpurun_out/r2k_thin6_raw.csv
ncu --set full --clock-control none --import-source on -k regex:slos_tile_kernel -c 24 -f -o gpurun_out/r2k_tile python tools/profile_slos.py 12 24 1 >> gpurun_out/r2k_prof.log 2>&1
ncu -i gpurun_out/r2k_tile.ncu-rep --page raw --csv > gpurun_out/r2k_tile_raw.csv
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2k_ref.json 2>/dev/null; tail -n1 gpurun_out/r2k_ref.json | cut -c1-300
ls gpurun_out | grep r2k
