# one ncu --set full capture of the table and tile-streaming kernels (nohits and R inputs)
tag=${1:-r2}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vy_decode_stream2|vy_decode_table' -s 20 -c 2 -o gpurun_out/${tag}_prof python tools/kernel_times.py coco608_b64 > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}_prof*
