#!/bin/bash
# Run under gpurun on ONE B200: ncu --set full of the shipped K1 kernel of every SF on an 8 GiB batch (the bench's launch
# size), the stream kernels, and the launch list of the bench command.  Outputs land in gpurun_out/; tools/make_traffic.py
# turns them into profiles/k1_traffic.json + profiles/r2_k1_sf*.txt here.
set -u
mkdir -p gpurun_out
for sf in 7 8 9 10 11 12; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k1_(sf7|group|sf10|rows)' -s 2 -c 1 \
      -o gpurun_out/r2_k1_sf$sf -f python tools/k1_ab.py --sf $sf --gib 8 --reps 1 --no-parity > gpurun_out/ncu_k1_sf$sf.log 2>&1
  tail -1 gpurun_out/ncu_k1_sf$sf.log | cut -c1-200
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'rx_warp' -c 1 -o gpurun_out/r2_rx_warp_final -f \
    python tools/rx_profile.py --sf 7 --streams 4096 --reps 1 > gpurun_out/ncu_rx_warp.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k1_|rx_|k8_|sc16_|sc8_|chan_' -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-config4 --no-cpu > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err
tail -2 gpurun_out/r2_launches_bench.csv | cut -c1-300
