for defs in "" "MOTIF_CORR_NOCOMPUTE"; do
  echo "== '$defs'"
  MOTIF_DEFINES="$defs" python -m motif_b200.build --force > /dev/null
  ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_registers,launch__waves_per_multiprocessor --clock-control none -k regex:corr_kernel --csv --log-file gpurun_out/corr_ab.csv python tools/run_corr.py > /dev/null 2>&1
  python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/corr_ab.csv')) if len(r)>10]
h=rows[0]
seen={}
for r in rows[1:]:
    seen.setdefault((r[h.index("ID")],r[h.index("Grid Size")]),[]).append((r[h.index("Metric Name")][:28],r[h.index("Metric Value")]))
for k,v in list(seen.items())[::3]:
    print(k[1], v)
PY
done
