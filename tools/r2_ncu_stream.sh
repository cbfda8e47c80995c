# one ncu --set full capture of the tile-streaming kernel (R inputs)
tag=${1:-r2}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'vy_decode_stream2' -s 10 -c 1 -o gpurun_out/${tag}_prof python tools/kernel_times.py coco608_b64 > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}_prof*
