set -x
timeout 200 python tools/conv_bench.py 8 > gpurun_out/t2_conv.log 2>&1
timeout 200 python tools/quick_time.py > gpurun_out/t2_quick.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/t2_launches.csv python bench.py --steps 10 --warmup 5 --no-cpu --no-e2e > gpurun_out/t2_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'vy_decode_stream|vy_decode_sample|vy_nms_finalize' -s 12 -c 3 -o gpurun_out/t2_prof python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/t2_p.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vy_fusion_conv -c 1 -o gpurun_out/t2_conv_prof python tools/conv_bench.py 8 one > gpurun_out/t2_cp.log 2>&1
cat gpurun_out/t2_conv.log gpurun_out/t2_quick.log
