timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/t16_grid.csv python tools/stress_time.py 4 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/t16_exh.csv python tools/stress_time.py 4 exh > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/t16_grid.csv","gpurun_out/t16_exh.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    agg={}
    for r in rows:
        k=(r[4][:60], r[12]); agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[14].replace(",",""))
    print(f)
    for (name,met),(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
        print("  %-62s %-28s n=%3d total=%.3g" % (name,met,n,v))
PY
