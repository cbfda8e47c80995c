# one GPU call: parity tests, bench line, ncu launch list + full capture of the hot kernels
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-run}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 10 --warmup 5 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'vy_decode_stream|vy_decode_sample|vy_nms_finalize|vy_decode_select' -s 16 -c 4 -o gpurun_out/${tag}_prof python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls -la gpurun_out/ | tail -5
